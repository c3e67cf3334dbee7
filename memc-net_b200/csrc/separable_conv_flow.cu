// separable_conv_flow.cu -- SeparableConvFlow: the flow a pair of separable filters encodes (their centroids),
//     flow_y = sum_k k * input2[b,k,h,w] / sum_k input2[b,k,h,w] - (fs - 1) / 2,   flow_x likewise from input3,
// on the valid region (H - fs + 1) x (W - fs + 1); a filter whose taps sum to exactly 0 gives -2000.  input1 only carries the
// frame size.  Backward: d flow / d tap k = (k / sum - centroid_sum / sum^2) * gradflow; nothing where the sum is 0.
//
// Semantics: reference my_package/src/my_lib_kernel.cu:19-83 (forward), :85-162 (backward), launchers :164-283; CPU twin
// my_lib.c:13-249, which divides by |sum| where the CUDA source divides by the signed sum -- this file follows the CUDA
// source (the two agree for filters with a positive sum, where the oracle is pinned).  The CUDA backward ASSIGNS
// gradinput2 but ACCUMULATES into gradinput3 (`=` at :131, `+=` at :155); kept under the reference contract, while
// MEMC_B200_OVERWRITE writes every element of both.  No Python class or caller in the reference.
#include "memc_common.cuh"

namespace memc {

namespace {

constexpr int BX = 32, BY = 8;

struct ScfArgs {
    int B, H, W, fs;        // H, W: the frame (input1); the op runs on (H - fs + 1) x (W - fs + 1)
    View vert, horiz, flow; // input2, input3 [B,fs,Ho,Wo]; flow_output / gradflow_output [B,2,Ho,Wo]
    View gv, gh;            // gradinput2, gradinput3
    const float* vertp;
    const float* horizp;
    float* flowp;           // fwd: written; bwd: gradflow_output (read)
    float* gvp;
    float* ghp;
};

// centroid sum and tap sum of one filter at one pixel
__device__ __forceinline__ void scf_sums(const float* f, int64_t stride_c, int fs, float& cen, float& sum) {
    cen = 0.0f; sum = 0.0f;
    for (int k = 0; k < fs; ++k) {
        const float t = __ldg(f + k * stride_c);
        cen += (float)k * t;
        sum += t;
    }
}

__global__ void __launch_bounds__(BX* BY) scf_fwd_kernel(const ScfArgs p) {
    const int w = blockIdx.x * BX + threadIdx.x, h = blockIdx.y * BY + threadIdx.y, b = blockIdx.z;
    if (w >= p.W - p.fs + 1 || h >= p.H - p.fs + 1) return;
    const double half = ((double)(float)p.fs - 1.0) / 2.0;  // "((float)(filter_size)-1.0)/2.0" is double arithmetic (:64)
    float cen, sum;
    scf_sums(p.vertp + b * p.vert.b + (int64_t)h * p.vert.h + w, p.vert.c, p.fs, cen, sum);
    float* o = p.flowp + b * p.flow.b + (int64_t)h * p.flow.h + w;
    o[p.flow.c] = fabsf(sum) > 0.0f ? (float)((double)(cen / sum) - half) : -2000.0f;
    scf_sums(p.horizp + b * p.horiz.b + (int64_t)h * p.horiz.h + w, p.horiz.c, p.fs, cen, sum);
    o[0] = fabsf(sum) > 0.0f ? (float)((double)(cen / sum) - half) : -2000.0f;
}

template <bool OVERWRITE>
__global__ void __launch_bounds__(BX* BY) scf_bwd_kernel(const ScfArgs p) {
    const int w = blockIdx.x * BX + threadIdx.x, h = blockIdx.y * BY + threadIdx.y, b = blockIdx.z;
    if (w >= p.W - p.fs + 1 || h >= p.H - p.fs + 1) return;
    const float* go = p.flowp + b * p.flow.b + (int64_t)h * p.flow.h + w;
    float cen, sum;
    scf_sums(p.vertp + b * p.vert.b + (int64_t)h * p.vert.h + w, p.vert.c, p.fs, cen, sum);
    float* gv = p.gvp + b * p.gv.b + (int64_t)h * p.gv.h + w;
    if (fabsf(sum) > 0.0f) {
        const float g = __ldg(go + p.flow.c), off = cen / (sum * sum);
        for (int k = 0; k < p.fs; ++k) gv[k * p.gv.c] = g * ((float)k / sum - off);  // assigned (:131)
    } else if (OVERWRITE) {
        for (int k = 0; k < p.fs; ++k) gv[k * p.gv.c] = 0.0f;
    }
    scf_sums(p.horizp + b * p.horiz.b + (int64_t)h * p.horiz.h + w, p.horiz.c, p.fs, cen, sum);
    float* gh = p.ghp + b * p.gh.b + (int64_t)h * p.gh.h + w;
    if (fabsf(sum) > 0.0f) {
        const float g = __ldg(go), off = cen / (sum * sum);
        for (int k = 0; k < p.fs; ++k) {
            const float v = g * ((float)k / sum - off);
            gh[k * p.gh.c] = OVERWRITE ? v : gh[k * p.gh.c] + v;  // accumulated by the reference (:155)
        }
    } else if (OVERWRITE) {
        for (int k = 0; k < p.fs; ++k) gh[k * p.gh.c] = 0.0f;
    }
}

int scf_launch(cudaStream_t stream, const ScfArgs& a, bool backward, int flags) {
    if (a.B <= 0 || a.fs <= 0 || a.H - a.fs + 1 <= 0 || a.W - a.fs + 1 <= 0) return 0;
    if (a.B > 65535) return -1;
    DeviceGuard guard(a.vertp);
    if (!guard.ok) return -1;
    dim3 block(BX, BY, 1), grid((a.W - a.fs + 1 + BX - 1) / BX, (a.H - a.fs + 1 + BY - 1) / BY, a.B);
    if (!backward) scf_fwd_kernel<<<grid, block, 0, stream>>>(a);
    else if (flags & MEMC_B200_OVERWRITE) scf_bwd_kernel<true><<<grid, block, 0, stream>>>(a);
    else scf_bwd_kernel<false><<<grid, block, 0, stream>>>(a);
    count_launch();
    return check_launch(backward ? "SeparableConvFlow backward" : "SeparableConvFlow forward");
}

}  // namespace

}  // namespace memc

using namespace memc;

extern "C" int memc_b200_separable_conv_flow_forward(
    memc_stream_t stream, int batch, int h, int w, int filter_size,
    memc_strides s_vert, memc_strides s_horiz, memc_strides s_flow,
    const float* vertical, const float* horizontal, float* flow_output, int flags) {
    ScfArgs a{};
    a.B = batch; a.H = h; a.W = w; a.fs = filter_size;
    a.vert = mk_view(s_vert); a.horiz = mk_view(s_horiz); a.flow = mk_view(s_flow);
    a.vertp = vertical; a.horizp = horizontal; a.flowp = flow_output;
    return scf_launch(stream, a, false, flags);
}

extern "C" int memc_b200_separable_conv_flow_backward(
    memc_stream_t stream, int batch, int h, int w, int filter_size,
    memc_strides s_vert, memc_strides s_horiz, memc_strides s_gflow, memc_strides s_gvert, memc_strides s_ghoriz,
    const float* vertical, const float* horizontal, const float* gradflow_output, float* gradvertical, float* gradhorizontal,
    int flags) {
    ScfArgs a{};
    a.B = batch; a.H = h; a.W = w; a.fs = filter_size;
    a.vert = mk_view(s_vert); a.horiz = mk_view(s_horiz); a.flow = mk_view(s_gflow); a.gv = mk_view(s_gvert); a.gh = mk_view(s_ghoriz);
    a.vertp = vertical; a.horizp = horizontal; a.flowp = const_cast<float*>(gradflow_output); a.gvp = gradvertical; a.ghp = gradhorizontal;
    return scf_launch(stream, a, true, flags);
}

// Reference-named launchers (my_lib_kernel.h:6-35); gradinput2 / gradinput3 use input2's / input3's strides, gradinput1 is not
// touched (the frame has no gradient here).
extern "C" int SeparableConvFlowLayer_gpu_forward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int channel, const int batch, const int filter_size,
    const int i1b, const int i1c, const int i1h, const int i1w,
    const int i2b, const int i2c, const int i2h, const int i2w,
    const int i3b, const int i3c, const int i3h, const int i3w,
    const int fb, const int fc, const int fh, const int fw,
    const float* input1, const float* input2, const float* input3, float* flow_output) {
    (void)nElement; (void)channel; (void)i1b; (void)i1c; (void)i1h; (void)i1w; (void)input1;
    if (i2w != 1 || i3w != 1 || fw != 1) return -1;
    ScfArgs a{};
    a.B = batch; a.H = h; a.W = w; a.fs = filter_size;
    a.vert = mk_view(i2b, i2c, i2h); a.horiz = mk_view(i3b, i3c, i3h); a.flow = mk_view(fb, fc, fh);
    a.vertp = input2; a.horizp = input3; a.flowp = flow_output;
    return scf_launch(stream, a, false, 0);
}

extern "C" int SeparableConvFlowLayer_gpu_backward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int channel, const int batch, const int filter_size,
    const int i1b, const int i1c, const int i1h, const int i1w,
    const int i2b, const int i2c, const int i2h, const int i2w,
    const int i3b, const int i3c, const int i3h, const int i3w,
    const int fb, const int fc, const int fh, const int fw,
    const float* input1, const float* input2, const float* input3, const float* gradflow_output,
    float* gradinput1, float* gradinput2, float* gradinput3) {
    (void)nElement; (void)channel; (void)i1b; (void)i1c; (void)i1h; (void)i1w; (void)input1; (void)gradinput1;
    if (i2w != 1 || i3w != 1 || fw != 1) return -1;
    ScfArgs a{};
    a.B = batch; a.H = h; a.W = w; a.fs = filter_size;
    a.vert = mk_view(i2b, i2c, i2h); a.horiz = mk_view(i3b, i3c, i3h); a.flow = mk_view(fb, fc, fh);
    a.gv = a.vert; a.gh = a.horiz;
    a.vertp = input2; a.horizp = input3; a.flowp = const_cast<float*>(gradflow_output); a.gvp = gradinput2; a.ghp = gradinput3;
    return scf_launch(stream, a, true, 0);
}
