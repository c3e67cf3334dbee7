/*
 * memc_b200.h -- C ABI of libmemc_b200.so: B200-native (sm_100a) kernels for MEMC-Net's
 * per-pixel motion-compensation hot path.
 *
 * Two groups of entry points:
 *
 *  (1) REFERENCE-NAMED LAUNCHERS.  Same names, argument order and error convention as the
 *      extern "C" launchers the reference declares in my_package/src/my_lib_kernel.h and
 *      calls from its THC glue my_package/src/my_lib_cuda.c.  A caller written against the
 *      reference's header (the cffi/THC glue, or a ctypes stub) links
 *      against libmemc_b200.so instead of my_lib_kernel.o with no source change.
 *      Contract (identical to the reference):
 *        - all tensors fp32, 4-D NCHW, w-stride 1; strides are in ELEMENTS (int);
 *        - the caller owns every buffer and ZERO-FILLS every output / gradient buffer
 *          before the call; the library allocates nothing and frees nothing;
 *        - work is enqueued asynchronously on `stream`; no synchronisation; the launch goes to the device that
 *          owns the operands (made current for the call and restored), `stream` must belong to it;
 *        - returns 0 on success, -1 on a launch failure or a layout the kernels cannot
 *          take (w-stride != 1);  `nElement` is accepted and ignored, as in the reference.
 *      Beyond the reference: batch/channel offsets are computed in 64 bits, so tensors
 *      above 2^31 elements work as long as each stride fits an int.
 *
 *  (2) memc_b200_* EXTENDED ENTRY POINTS.  64-bit strides plus a `flags` word:
 *        MEMC_B200_OVERWRITE  the library itself produces every element of every output
 *                             (zero-filling what it scatters into), so the caller may pass
 *                             uninitialised buffers and skip its memsets;
 *        MEMC_B200_NO_FAST    force the generic (non-TMA) kernels (used by the tests to
 *                             cross-check the fast path);
 *        MEMC_B200_FLOAT_ACCUM  FilterInterpolation backward: accumulate gradinput1 with fp32 shared
 *                             atomics instead of the per-tile fixed point (slower; keeps fp32
 *                             RELATIVE precision for gradients whose magnitude varies by many
 *                             orders inside one 32x8 tile -- the fixed point guarantees an
 *                             absolute error of ~2^-22 x the tile's largest contribution);
 *        MEMC_B200_NO_ZERO    with OVERWRITE: the caller has already zero-filled the
 *                             scatter targets (gradinput1 / count+output); lets bench.py
 *                             time the kernel apart from the memset.
 *      The Python autograd Functions in memc-net_b200/my_package/functions use these.
 *
 * cudaStream_t is passed as an opaque pointer so that this header needs no CUDA include.
 */
#ifndef MEMC_B200_H
#define MEMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef MEMC_B200_STREAM_T
#define MEMC_B200_STREAM_T
typedef void *memc_stream_t; /* a cudaStream_t */
#endif

#if defined(__GNUC__)
#define MEMC_B200_API __attribute__((visibility("default")))
#else
#define MEMC_B200_API
#endif

#define MEMC_B200_OVERWRITE 1
#define MEMC_B200_NO_FAST 2
#define MEMC_B200_NO_ZERO 4
#define MEMC_B200_FLOAT_ACCUM 8
/* bits 16..23: kernel variant, 0 = production.  Non-zero values keep earlier kernels reachable for A/B
 * measurements and for the tests' cross-checks (FilterInterpolation forward, C <= 4: 1 row segments, 2 patches with
 * a TMA-staged output; C > 4: 1 generic kernel, 2 the round-1 patch kernel; backward: 1 the round-1
 * one-pixel-per-lane kernel, 2-4 other tile shapes; FlowProjection forward: 1 frame-by-frame launches instead of the
 * persistent pipeline, 2 the pipeline without its L2 eviction policies; DepthFlowProjection forward: 1 the generic
 * scatter inside the frame driver).  Nothing in the library reads environment variables. */
#define MEMC_B200_VARIANT(n) (((n) & 0xff) << 16)

/* ---- library info ------------------------------------------------------------------ */
MEMC_B200_API int memc_b200_abi_version(void);          /* bumped on any signature change            */
MEMC_B200_API const char *memc_b200_build_info(void);   /* "sm_100a nvcc <ver> ..."                   */
/* number of kernel launches (incl. memsets) the library has issued since load, for
 * bench.py's gpu_launches accounting */
MEMC_B200_API unsigned long long memc_b200_launch_count(void);
/* The only memory the library allocates is stream-ordered scratch from a private per-device pool (FlowProjection's
 * accumulators / occupancy masks, the unfused blend fallback); up to 1 GiB stays cached between calls.  This hands
 * all of it back to the driver.  0 = ok. */
MEMC_B200_API int memc_b200_scratch_trim(void);

/* ====================================================================================
 * (1) reference-named launchers
 * ==================================================================================== */

/* replaces my_lib_kernel.h:133-144 (called from my_lib_cuda.c:651) */
MEMC_B200_API int FilterInterpolationLayer_gpu_forward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch, const int filter_size,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input2_b_stride, const int input2_c_stride, const int input2_h_stride, const int input2_w_stride,
    const int input3_b_stride, const int input3_c_stride, const int input3_h_stride, const int input3_w_stride,
    const float *input1, const float *input2, const float *input3, float *output);

/* replaces my_lib_kernel.h:146-158 (called from my_lib_cuda.c:729) */
MEMC_B200_API int FilterInterpolationLayer_gpu_backward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch, const int filter_size,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input2_b_stride, const int input2_c_stride, const int input2_h_stride, const int input2_w_stride,
    const int input3_b_stride, const int input3_c_stride, const int input3_h_stride, const int input3_w_stride,
    const float *input1, const float *input2, const float *input3,
    const float *gradoutput, float *gradinput1, float *gradinput2, float *gradinput3);

/* replaces my_lib_kernel.h:161-171 (called from my_lib_cuda.c:785): scatter -> average ->
 * (fillhole ? fill-hole : nothing) */
MEMC_B200_API int FlowProjection_gpu_forward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch, const int fillhole,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int count_b_stride, const int count_c_stride, const int count_h_stride, const int count_w_stride,
    const float *input1, float *count, float *output);

/* replaces my_lib_kernel.h:173-187 (called from my_lib_cuda.c:839) */
MEMC_B200_API int FlowProjection_gpu_backward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int count_b_stride, const int count_c_stride, const int count_h_stride, const int count_w_stride,
    const float *input1, const float *count, const float *gradoutput, float *gradinput1);

/* replaces my_lib_kernel.h:189-200 (called from my_lib_cuda.c:898): FlowProjection with a per-source weight
 * (input2 [B,1,H,W], e.g. an inverse depth): scatter -w*flow and w -> divide where the accumulated weight > 0 ->
 * (fillhole ? fill-hole : nothing).  SURVEY section 8(f) rank 4; the reference ships this C side but no Python class. */
MEMC_B200_API int DepthFlowProjection_gpu_forward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch, const int fillhole,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input2_b_stride, const int input2_c_stride, const int input2_h_stride, const int input2_w_stride,
    const int count_b_stride, const int count_c_stride, const int count_h_stride, const int count_w_stride,
    const float *input1, const float *input2, float *count, float *output);

/* replaces my_lib_kernel.h:202-220 (called from my_lib_cuda.c:963); `output` is the forward's result */
MEMC_B200_API int DepthFlowProjection_gpu_backward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input2_b_stride, const int input2_c_stride, const int input2_h_stride, const int input2_w_stride,
    const int count_b_stride, const int count_c_stride, const int count_h_stride, const int count_w_stride,
    const float *input1, const float *input2, const float *count, const float *output, const float *gradoutput,
    float *gradinput1, float *gradinput2);

/* replaces my_lib_kernel.h:222-236 (called from my_lib_cuda.c:1041): FlowProjection in which a source votes only if its
 * brightness-constancy error mean_c|input2[h,w] - input3[(h,w) + 2 flow]| + 1e-8 is <= threshhold; the error is splatted
 * and averaged into `weight` [B,1,H,W] as well.  input2 / input3 are [B,3,H,W].  SURVEY section 8(f) rank 4. */
MEMC_B200_API int WeightedFlowProjection_gpu_forward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch, const int fillhole, const float threshhold,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input2_b_stride, const int input2_c_stride, const int input2_h_stride, const int input2_w_stride,
    const int input3_b_stride, const int input3_c_stride, const int input3_h_stride, const int input3_w_stride,
    const int count_b_stride, const int count_c_stride, const int count_h_stride, const int count_w_stride,
    const int weight_b_stride, const int weight_c_stride, const int weight_h_stride, const int weight_w_stride,
    const float *input1, const float *input2, const float *input3, float *count, float *weight, float *output);

/* replaces my_lib_kernel.h:238-255 (called from my_lib_cuda.c:1122) */
MEMC_B200_API int WeightedFlowProjection_gpu_backward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch, const float threshhold,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input2_b_stride, const int input2_c_stride, const int input2_h_stride, const int input2_w_stride,
    const int input3_b_stride, const int input3_c_stride, const int input3_h_stride, const int input3_w_stride,
    const int count_b_stride, const int count_c_stride, const int count_h_stride, const int count_w_stride,
    const int weight_b_stride, const int weight_c_stride, const int weight_h_stride, const int weight_w_stride,
    const float *input1, const float *input2, const float *input3, const float *count, const float *weight,
    const float *gradoutput, float *gradinput1);

/* replaces my_lib_kernel.h:257-270 (called from my_lib_cuda.c:1218): a per-pixel matching confidence, output [B,1,H,W] =
 * (1 - err / lambda_e)^2 with err = the mean absolute difference over the 3x3 neighbourhood and the channels between input1
 * around the pixel and input2 bilinearly sampled around pixel + flow (input3); 1e-4 where the target is outside the frame.
 * Nw must be 3; lambda_v is unused (as in the reference).  No Python class or caller in the reference. */
MEMC_B200_API int WeightLayer_gpu_forward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input2_b_stride, const int input2_c_stride, const int input2_h_stride, const int input2_w_stride,
    const int input3_b_stride, const int input3_c_stride, const int input3_h_stride, const int input3_w_stride,
    const int output_b_stride, const int output_c_stride, const int output_h_stride, const int output_w_stride,
    const float *input1, const float *input2, const float *input3, float *output,
    float lambda_e, float lambda_v, float Nw);

/* replaces my_lib_kernel.h:271-286 (called from my_lib_cuda.c:1318); `output` is the forward's result */
MEMC_B200_API int WeightLayer_gpu_backward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input2_b_stride, const int input2_c_stride, const int input2_h_stride, const int input2_w_stride,
    const int input3_b_stride, const int input3_c_stride, const int input3_h_stride, const int input3_w_stride,
    const int output_b_stride, const int output_c_stride, const int output_h_stride, const int output_w_stride,
    const float *input1, const float *input2, const float *input3, const float *output, const float *gradoutput,
    float *gradinput1, float *gradinput2, float *gradinput3,
    float lambda_e, float lambda_v, float Nw);

/* replaces my_lib_kernel.h:6-19 (called from my_lib_cuda.c:80): the flow a pair of separable filters encodes -- centroid of
 * input2's taps minus (fs-1)/2 -> channel 1, of input3's -> channel 0, on the (h-fs+1) x (w-fs+1) valid region; -2000 where
 * the taps sum to 0.  input1 only carries the frame size.  No Python class or caller in the reference. */
MEMC_B200_API int SeparableConvFlowLayer_gpu_forward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch, const int filter_size,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input2_b_stride, const int input2_c_stride, const int input2_h_stride, const int input2_w_stride,
    const int input3_b_stride, const int input3_c_stride, const int input3_h_stride, const int input3_w_stride,
    const int flow_output_b_stride, const int flow_output_c_stride, const int flow_output_h_stride, const int flow_output_w_stride,
    const float *input1, const float *input2, const float *input3, float *flow_output);

/* replaces my_lib_kernel.h:21-35: gradinput2 is ASSIGNED, gradinput3 accumulated (my_lib_kernel.cu:131, 155);
 * gradinput1 is not touched */
MEMC_B200_API int SeparableConvFlowLayer_gpu_backward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch, const int filter_size,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input2_b_stride, const int input2_c_stride, const int input2_h_stride, const int input2_w_stride,
    const int input3_b_stride, const int input3_c_stride, const int input3_h_stride, const int input3_w_stride,
    const int flow_output_b_stride, const int flow_output_c_stride, const int flow_output_h_stride, const int flow_output_w_stride,
    const float *input1, const float *input2, const float *input3, const float *gradflow_output,
    float *gradinput1, float *gradinput2, float *gradinput3);

/* ---- the 4x4 "pixel splat" family (no Python class or caller in the reference) ----
 * Every source pixel lands at (w, h) + flow / 2 and spreads over the 4x4 cells around it with the window weight
 * (1 - ((beta - m)^2 + (alpha - n)^2) / (2 sigma_d^2))^2:  PixelValue adds flow_weight * weight * input1[c], PixelWeight
 * flow_weight * weight, ReliableWeight the weight itself.  Prowindow must be 2 and tao_r is unused (as in the reference).
 * replaces my_lib_kernel.h:297-308 (called from my_lib_cuda.c:1413) */
MEMC_B200_API int PixelValueLayer_gpu_forward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input3_b_stride, const int input3_c_stride, const int input3_h_stride, const int input3_w_stride,
    const int flow_weights_b_stride, const int flow_weights_c_stride, const int flow_weights_h_stride, const int flow_weights_w_stride,
    const int output_b_stride, const int output_c_stride, const int output_h_stride, const int output_w_stride,
    const float *input1, const float *input3, const float *flow_weights, float *output,
    float sigma_d, float tao_r, float Prowindow);

/* replaces my_lib_kernel.h:309-322 (called from my_lib_cuda.c:1510) */
MEMC_B200_API int PixelValueLayer_gpu_backward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input3_b_stride, const int input3_c_stride, const int input3_h_stride, const int input3_w_stride,
    const int flow_weights_b_stride, const int flow_weights_c_stride, const int flow_weights_h_stride, const int flow_weights_w_stride,
    const int output_b_stride, const int output_c_stride, const int output_h_stride, const int output_w_stride,
    const float *input1, const float *input3, const float *flow_weights,
    const float *gradoutput, float *gradinput1, float *gradinput3, float *gradflow_weights,
    float sigma_d, float tao_r, float Prowindow);

/* replaces my_lib_kernel.h:323-334 (called from my_lib_cuda.c:1604) */
MEMC_B200_API int PixelWeightLayer_gpu_forward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int batch,
    const int input3_b_stride, const int input3_c_stride, const int input3_h_stride, const int input3_w_stride,
    const int flow_weights_b_stride, const int flow_weights_c_stride, const int flow_weights_h_stride, const int flow_weights_w_stride,
    const int output_b_stride, const int output_c_stride, const int output_h_stride, const int output_w_stride,
    const float *input3, const float *flow_weights, float *output,
    float sigma_d, float tao_r, float Prowindow);

/* replaces my_lib_kernel.h:335-349 (called from my_lib_cuda.c:1697): cells whose forward output is below threshhold give no gradient */
MEMC_B200_API int PixelWeightLayer_gpu_backward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int batch,
    const int input3_b_stride, const int input3_c_stride, const int input3_h_stride, const int input3_w_stride,
    const int flow_weights_b_stride, const int flow_weights_c_stride, const int flow_weights_h_stride, const int flow_weights_w_stride,
    const int output_b_stride, const int output_c_stride, const int output_h_stride, const int output_w_stride,
    const float *input3, const float *flow_weights, const float *output,
    const float *gradoutput, float *gradinput3, float *gradflow_weights,
    float threshhold, float sigma_d, float tao_r, float Prowindow);

/* replaces my_lib_kernel.h:350-361 (called from my_lib_cuda.c:1809) */
MEMC_B200_API int ReliableWeightLayer_gpu_forward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int batch,
    const int input3_b_stride, const int input3_c_stride, const int input3_h_stride, const int input3_w_stride,
    const int output_b_stride, const int output_c_stride, const int output_h_stride, const int output_w_stride,
    const float *input3, float *output,
    float sigma_d, float tao_r, float Prowindow);

/* replaces my_lib_kernel.h:362-377 (called from my_lib_cuda.c:1904) */
MEMC_B200_API int ReliableWeightLayer_gpu_backward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int batch,
    const int input3_b_stride, const int input3_c_stride, const int input3_h_stride, const int input3_w_stride,
    const int output_b_stride, const int output_c_stride, const int output_h_stride, const int output_w_stride,
    const float *input3, const float *output, const float *gradoutput, float *gradinput3,
    float threshhold, float sigma_d, float tao_r, float Prowindow);

/* replaces my_lib_kernel.h:67-81 (called from my_lib_cuda.c:402 and, for the Ch variant,
 * :519) */
MEMC_B200_API int InterpolationLayer_gpu_forward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input2_b_stride, const int input2_c_stride, const int input2_h_stride, const int input2_w_stride,
    const float *input1, const float *input2, float *output);

/* replaces my_lib_kernel.h:83-98 (called from my_lib_cuda.c:462, :579) */
MEMC_B200_API int InterpolationLayer_gpu_backward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input2_b_stride, const int input2_c_stride, const int input2_h_stride, const int input2_w_stride,
    const float *input1, const float *input2, const float *gradoutput,
    float *gradinput1, float *gradinput2);

/* replace my_lib_kernel.h:101-132: the reference's InterpolationCh kernels are a copy of
 * the Interpolation ones with no channel restriction; here they are the same code. */
MEMC_B200_API int InterpolationChLayer_gpu_forward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input2_b_stride, const int input2_c_stride, const int input2_h_stride, const int input2_w_stride,
    const float *input1, const float *input2, float *output);

MEMC_B200_API int InterpolationChLayer_gpu_backward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input2_b_stride, const int input2_c_stride, const int input2_h_stride, const int input2_w_stride,
    const float *input1, const float *input2, const float *gradoutput,
    float *gradinput1, float *gradinput2);

/* replaces my_lib_kernel.h:37-48 (called from my_lib_cuda.c:257).  w,h are the INPUT
 * image extent; the output/filter extent is (h-fs+1) x (w-fs+1). */
MEMC_B200_API int SeparableConvLayer_gpu_forward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch, const int filter_size,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input2_b_stride, const int input2_c_stride, const int input2_h_stride, const int input2_w_stride,
    const int input3_b_stride, const int input3_c_stride, const int input3_h_stride, const int input3_w_stride,
    const int output_b_stride, const int output_c_stride, const int output_h_stride, const int output_w_stride,
    const float *input1, const float *input2, const float *input3, float *output);

/* replaces my_lib_kernel.h:50-64 (called from my_lib_cuda.c:341) */
MEMC_B200_API int SeparableConvLayer_gpu_backward_kernel(
    memc_stream_t stream, const int nElement,
    const int w, const int h, const int channel, const int batch, const int filter_size,
    const int input1_b_stride, const int input1_c_stride, const int input1_h_stride, const int input1_w_stride,
    const int input2_b_stride, const int input2_c_stride, const int input2_h_stride, const int input2_w_stride,
    const int input3_b_stride, const int input3_c_stride, const int input3_h_stride, const int input3_w_stride,
    const int output_b_stride, const int output_c_stride, const int output_h_stride, const int output_w_stride,
    const float *input1, const float *input2, const float *input3,
    const float *gradoutput, float *gradinput1, float *gradinput2, float *gradinput3);

/* ====================================================================================
 * (2) extended entry points
 * ==================================================================================== */

/* strides of one NCHW tensor, in elements; the w-stride is implicitly 1 */
typedef struct memc_strides {
    int64_t b, c, h;
} memc_strides;

MEMC_B200_API int memc_b200_filter_interpolation_forward(
    memc_stream_t stream, int batch, int channel, int h, int w, int filter_size,
    memc_strides s_in1, memc_strides s_flow, memc_strides s_filter, memc_strides s_out,
    const float *input1, const float *flow, const float *filter, float *output, int flags);

MEMC_B200_API int memc_b200_filter_interpolation_backward(
    memc_stream_t stream, int batch, int channel, int h, int w, int filter_size,
    memc_strides s_in1, memc_strides s_flow, memc_strides s_filter, memc_strides s_gout,
    memc_strides s_gi1, memc_strides s_gi2, memc_strides s_gi3,
    const float *input1, const float *flow, const float *filter, const float *gradoutput,
    float *gradinput1, float *gradinput2, float *gradinput3, int flags);

/* Fused call site of the reference's networks (networks/MEMC_Net.py:258-264, MEMC_Net_star.py:272-278,
 * `FilterInterpolate`):
 *     output = occlusion_0 * FilterInterpolation(input1_0, flow_0, filter_0)
 *            + occlusion_1 * FilterInterpolation(input1_1, flow_1, filter_1)
 * occlusion_k is [B,1,H,W] (broadcast over the channels); passing NULL for BOTH maps means the constant 0.5, i.e.
 * the plain mean `warp0 / 2 + warp1 / 2` of networks/MEMC_Net_s.py:260-264 (exact: halving is exact in fp32);
 * every element of `output` is written
 * (OVERWRITE semantics whatever `flags` says); bit-identical to the composition of
 * memc_b200_filter_interpolation_forward and the fp32 blend.  fs == 4, C == 3 take one fused
 * kernel; anything else is composed inside the library (two warps into stream-ordered scratch). */
MEMC_B200_API int memc_b200_filter_interpolation_blend_forward(
    memc_stream_t stream, int batch, int channel, int h, int w, int filter_size,
    memc_strides s_in1_0, memc_strides s_flow_0, memc_strides s_filter_0,
    memc_strides s_in1_1, memc_strides s_flow_1, memc_strides s_filter_1,
    memc_strides s_occ_0, memc_strides s_occ_1, memc_strides s_out,
    const float *input1_0, const float *flow_0, const float *filter_0,
    const float *input1_1, const float *flow_1, const float *filter_1,
    const float *occlusion_0, const float *occlusion_1, float *output, int flags);

/* Two images warped with ONE flow and ONE filter in one pass:
 *     output_a = FilterInterpolation(input_a, flow, filter),   output_b = FilterInterpolation(input_b, flow, filter)
 * -- the RGB frame and its context features in networks/MEMC_Net_star.py:272-285, which call the op twice per reference
 * with identical offset / filter.  flow, filter, geometry and bounding box are fetched / computed once (fs == 4, even H,
 * dense filter batch stride; otherwise, and with MEMC_B200_NO_FAST, the two plain calls are made inside the library).
 * Every element of both outputs is written. */
MEMC_B200_API int memc_b200_filter_interpolation_forward_pair(
    memc_stream_t stream, int batch, int channel_a, int channel_b, int h, int w, int filter_size,
    memc_strides s_in_a, memc_strides s_in_b, memc_strides s_flow, memc_strides s_filter,
    memc_strides s_out_a, memc_strides s_out_b,
    const float *input_a, const float *input_b, const float *flow, const float *filter,
    float *output_a, float *output_b, int flags);

MEMC_B200_API int memc_b200_flow_projection_forward(
    memc_stream_t stream, int batch, int h, int w, int fillhole,
    memc_strides s_flow, memc_strides s_count, memc_strides s_out,
    const float *flow, float *count, float *output, int flags);

MEMC_B200_API int memc_b200_flow_projection_backward(
    memc_stream_t stream, int batch, int h, int w,
    memc_strides s_flow, memc_strides s_count, memc_strides s_gout, memc_strides s_gi,
    const float *flow, const float *count, const float *gradoutput, float *gradinput, int flags);

MEMC_B200_API int memc_b200_depth_flow_projection_forward(
    memc_stream_t stream, int batch, int h, int w, int fillhole,
    memc_strides s_flow, memc_strides s_depth, memc_strides s_count, memc_strides s_out,
    const float *flow, const float *depth, float *count, float *output, int flags);

MEMC_B200_API int memc_b200_depth_flow_projection_backward(
    memc_stream_t stream, int batch, int h, int w,
    memc_strides s_flow, memc_strides s_depth, memc_strides s_count, memc_strides s_out, memc_strides s_gout,
    memc_strides s_gi1, memc_strides s_gi2,
    const float *flow, const float *depth, const float *count, const float *output, const float *gradoutput,
    float *gradinput1, float *gradinput2, int flags);

MEMC_B200_API int memc_b200_weighted_flow_projection_forward(
    memc_stream_t stream, int batch, int h, int w, int fillhole, float threshold,
    memc_strides s_flow, memc_strides s_frame0, memc_strides s_frame1, memc_strides s_count, memc_strides s_weight,
    memc_strides s_out,
    const float *flow, const float *frame0, const float *frame1, float *count, float *weight, float *output, int flags);

MEMC_B200_API int memc_b200_weighted_flow_projection_backward(
    memc_stream_t stream, int batch, int h, int w, float threshold,
    memc_strides s_flow, memc_strides s_frame0, memc_strides s_frame1, memc_strides s_count, memc_strides s_gout,
    memc_strides s_gi,
    const float *flow, const float *frame0, const float *frame1, const float *count, const float *gradoutput,
    float *gradinput, int flags);

MEMC_B200_API int memc_b200_weight_layer_forward(
    memc_stream_t stream, int batch, int channel, int h, int w, float lambda_e, float Nw,
    memc_strides s_in1, memc_strides s_in2, memc_strides s_flow, memc_strides s_out,
    const float *input1, const float *input2, const float *flow, float *output, int flags);

/* gradients share their inputs' strides, gradoutput the output's */
MEMC_B200_API int memc_b200_weight_layer_backward(
    memc_stream_t stream, int batch, int channel, int h, int w, float lambda_e, float Nw,
    memc_strides s_in1, memc_strides s_in2, memc_strides s_flow, memc_strides s_out,
    const float *input1, const float *input2, const float *flow, const float *output, const float *gradoutput,
    float *gradinput1, float *gradinput2, float *gradinput3, int flags);

MEMC_B200_API int memc_b200_separable_conv_flow_forward(
    memc_stream_t stream, int batch, int h, int w, int filter_size,
    memc_strides s_vert, memc_strides s_horiz, memc_strides s_flow,
    const float *vertical, const float *horizontal, float *flow_output, int flags);

MEMC_B200_API int memc_b200_separable_conv_flow_backward(
    memc_stream_t stream, int batch, int h, int w, int filter_size,
    memc_strides s_vert, memc_strides s_horiz, memc_strides s_gflow, memc_strides s_gvert, memc_strides s_ghoriz,
    const float *vertical, const float *horizontal, const float *gradflow_output, float *gradvertical, float *gradhorizontal,
    int flags);

MEMC_B200_API int memc_b200_pixel_value_forward(
    memc_stream_t stream, int batch, int channel, int h, int w, float sigma_d,
    memc_strides s_in1, memc_strides s_flow, memc_strides s_fw, memc_strides s_out,
    const float *input1, const float *flow, const float *flow_weights, float *output, int flags);

MEMC_B200_API int memc_b200_pixel_value_backward(
    memc_stream_t stream, int batch, int channel, int h, int w, float sigma_d,
    memc_strides s_in1, memc_strides s_flow, memc_strides s_fw, memc_strides s_gout, memc_strides s_gi1, memc_strides s_gi3,
    memc_strides s_gfw,
    const float *input1, const float *flow, const float *flow_weights, const float *gradoutput, float *gradinput1,
    float *gradinput3, float *gradflow_weights, int flags);

MEMC_B200_API int memc_b200_pixel_weight_forward(
    memc_stream_t stream, int batch, int h, int w, float sigma_d,
    memc_strides s_flow, memc_strides s_fw, memc_strides s_out,
    const float *flow, const float *flow_weights, float *output, int flags);

MEMC_B200_API int memc_b200_pixel_weight_backward(
    memc_stream_t stream, int batch, int h, int w, float sigma_d, float threshold,
    memc_strides s_flow, memc_strides s_fw, memc_strides s_out, memc_strides s_gout, memc_strides s_gi3, memc_strides s_gfw,
    const float *flow, const float *flow_weights, const float *output, const float *gradoutput, float *gradinput3,
    float *gradflow_weights, int flags);

MEMC_B200_API int memc_b200_reliable_weight_forward(
    memc_stream_t stream, int batch, int h, int w, float sigma_d,
    memc_strides s_flow, memc_strides s_out, const float *flow, float *output, int flags);

MEMC_B200_API int memc_b200_reliable_weight_backward(
    memc_stream_t stream, int batch, int h, int w, float sigma_d, float threshold,
    memc_strides s_flow, memc_strides s_out, memc_strides s_gout, memc_strides s_gi3,
    const float *flow, const float *output, const float *gradoutput, float *gradinput3, int flags);

MEMC_B200_API int memc_b200_interpolation_forward(
    memc_stream_t stream, int batch, int channel, int h, int w,
    memc_strides s_in1, memc_strides s_flow, memc_strides s_out,
    const float *input1, const float *flow, float *output, int flags);

MEMC_B200_API int memc_b200_interpolation_backward(
    memc_stream_t stream, int batch, int channel, int h, int w,
    memc_strides s_in1, memc_strides s_flow, memc_strides s_gout,
    memc_strides s_gi1, memc_strides s_gi2,
    const float *input1, const float *flow, const float *gradoutput,
    float *gradinput1, float *gradinput2, int flags);

MEMC_B200_API int memc_b200_separable_conv_forward(
    memc_stream_t stream, int batch, int channel, int h, int w, int filter_size,
    memc_strides s_in1, memc_strides s_vert, memc_strides s_horiz, memc_strides s_out,
    const float *input1, const float *vertical, const float *horizontal, float *output, int flags);

MEMC_B200_API int memc_b200_separable_conv_backward(
    memc_stream_t stream, int batch, int channel, int h, int w, int filter_size,
    memc_strides s_in1, memc_strides s_vert, memc_strides s_horiz, memc_strides s_gout,
    memc_strides s_gi1, memc_strides s_gi2, memc_strides s_gi3,
    const float *input1, const float *vertical, const float *horizontal, const float *gradoutput,
    float *gradinput1, float *gradinput2, float *gradinput3, int flags);

#ifdef __cplusplus
}
#endif
#endif /* MEMC_B200_H */
