// weighted_flow_projection.cu -- WeightedFlowProjection: FlowProjection in which a source
// pixel votes only if the two frames agree along its flow -- brightness-constancy error
//     e = mean_c |input2[b,c,h,w] - input3[b,c,y3,x3]| + 1e-8,   (x3, y3) = trunc(clamp((w,h) + 2 flow, 0, (W-1,H-1)))
// not above `threshhold` -- and in which the error itself is splatted and averaged into a third plane (`weight`).
//
// Semantics: reference my_package/src/my_lib_kernel.cu:2499-2618 (scatter), :2620-2657 (average of output AND weight),
// :2660-2762 (fill-hole: output only, FlowProjection's walks), :2764-2843 (backward: FlowProjection's gather behind the
// same gate), launchers :2845-3024; CPU twin my_lib.c:1879-2250 (no fill-hole).  FFI names my_lib_cuda.h:119-138; the
// reference has no Python class for it.
//
// count is integer valued, as in FlowProjection, so the average + occupancy-mask + mask fill-hole kernels of
// flow_projection_fast.cu are used unchanged (fp_frames_fast with this file's splat); the weight plane is divided by
// count in one more pass.  The splat is one source per thread with global reductions (16 per voting source).  A gated
// four-plane version of FlowProjection's shared-memory corner histogram was built and measured (B=16 x 1080p, 44 % of
// the sources voting): 1.86 ms against 1.75 ms for this kernel on a smooth field, 2.40 = 2.40 ms on the convergent one
// -- the L2 absorbs the reductions of neighbouring sources, and what made the legacy kernels take 32 ms on convergent
// flow was the fill-hole walk, which the occupancy masks remove -- so it was not kept.
#include "flow_projection.cuh"

namespace memc {

namespace {

constexpr int BX = 32, BY = 8;

struct WfpArgs {
    FpArgs f;          // flow / count / out (fwd: output, bwd: gradoutput) / gi
    View im0, im1;     // input2, input3 [B,3,H,W]
    const float* im0p;
    const float* im1p;
    View wgt;          // weight [B,1,H,W]
    float* wgtp;
    float threshold;
};

// the gate of a source pixel with a valid target (my_lib_kernel.cu:2563-2577): the error, in the reference's fp32 order
__device__ __forceinline__ float wfp_error(const WfpArgs& p, int b, int h, int w, float fx, float fy) {
    const int W = p.f.W, H = p.f.H;
    const int x3 = (int)fmaxf(fminf((float)w + 2.0f * fx, (float)W - 1.0f), 0.0f);
    const int y3 = (int)fmaxf(fminf((float)h + 2.0f * fy, (float)H - 1.0f), 0.0f);
    const float* a = p.im0p + b * p.im0.b + (int64_t)h * p.im0.h + w;
    const float* c = p.im1p + b * p.im1.b + (int64_t)y3 * p.im1.h + x3;
    float e = 0.0f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) e += fabsf(__ldg(a + ch * p.im0.c) - __ldg(c + ch * p.im1.c)) / 3.0f;
    return e + 1e-8f;
}

__device__ __forceinline__ bool wfp_valid(float x2, float y2, int W, int H) {
    return x2 >= 0.0f && y2 >= 0.0f && x2 <= (float)(W - 1) && y2 <= (float)(H - 1);
}

// frame b0 + blockIdx.z
__global__ void __launch_bounds__(BX* BY) wfp_scatter_kernel(const WfpArgs p, const int b0) {
    const int w = blockIdx.x * BX + threadIdx.x;
    const int h = blockIdx.y * BY + threadIdx.y;
    const int b = b0 + blockIdx.z;
    const FpArgs& f = p.f;
    if (w >= f.W || h >= f.H) return;
    const float* fl = f.flowp + b * f.flow.b + (int64_t)h * f.flow.h + w;
    const float fx = ldg_stream(fl);
    const float fy = ldg_stream(fl + f.flow.c);
    const float x2 = (float)w + fx, y2 = (float)h + fy;
    if (!wfp_valid(x2, y2, f.W, f.H)) return;
    const float e = wfp_error(p, b, h, w, fx, fy);
    if (!(e <= p.threshold)) return;  // only the flow vectors with a low brightness error vote (:2578)
    const int L = (int)x2, T = (int)y2;
    const int R = min(L + 1, f.W - 1), Bm = min(T + 1, f.H - 1);
    float* ox = f.outp + b * f.out.b;
    float* oy = ox + f.out.c;
    float* cn = f.countp + b * f.count.b;
    float* wg = p.wgtp + b * p.wgt.b;
    const int64_t oT = (int64_t)T * f.out.h, oB = (int64_t)Bm * f.out.h;
    const int64_t cT = (int64_t)T * f.count.h, cB = (int64_t)Bm * f.count.h;
    const int64_t wT = (int64_t)T * p.wgt.h, wB = (int64_t)Bm * p.wgt.h;
    red_add(ox + oT + L, -fx); red_add(ox + oT + R, -fx);
    red_add(ox + oB + L, -fx); red_add(ox + oB + R, -fx);
    red_add(oy + oT + L, -fy); red_add(oy + oT + R, -fy);
    red_add(oy + oB + L, -fy); red_add(oy + oB + R, -fy);
    red_add(cn + cT + L, 1.0f); red_add(cn + cT + R, 1.0f);
    red_add(cn + cB + L, 1.0f); red_add(cn + cB + R, 1.0f);
    red_add(wg + wT + L, e); red_add(wg + wT + R, e);
    red_add(wg + wB + L, e); red_add(wg + wB + R, e);
}

// weight /= count where count > 0 (my_lib_kernel.cu:2651-2654)
__global__ void __launch_bounds__(BX* BY) wfp_weight_average_kernel(const WfpArgs p) {
    const int w = blockIdx.x * BX + threadIdx.x;
    const int h = blockIdx.y * BY + threadIdx.y;
    const int b = blockIdx.z;
    if (w >= p.f.W || h >= p.f.H) return;
    const float c = p.f.countp[b * p.f.count.b + (int64_t)h * p.f.count.h + w];
    if (c > 0.0f) {
        float* g = p.wgtp + b * p.wgt.b + (int64_t)h * p.wgt.h + w;
        *g = *g / c;
    }
}

// gradinput1[ch] = - sum over the 4 cells of gradoutput[ch] / count, for the sources that voted (my_lib_kernel.cu:2806-2838)
template <bool OVERWRITE>
__global__ void __launch_bounds__(BX* BY, 6) wfp_bwd_kernel(const WfpArgs p) {
    const int w = blockIdx.x * BX + threadIdx.x;
    const int h = blockIdx.y * BY + threadIdx.y;
    const int b = blockIdx.z;
    const FpArgs& f = p.f;
    if (w >= f.W || h >= f.H) return;
    const float* fl = f.flowp + b * f.flow.b + (int64_t)h * f.flow.h + w;
    const float fx = ldg_stream(fl), fy = ldg_stream(fl + f.flow.c);
    float* gx = f.gip + b * f.gi.b + (int64_t)h * f.gi.h + w;
    float* gy = gx + f.gi.c;
    const float x2 = (float)w + fx, y2 = (float)h + fy;
    bool votes = wfp_valid(x2, y2, f.W, f.H);
    if (votes) votes = wfp_error(p, b, h, w, fx, fy) <= p.threshold;
    if (!votes) {
        if (OVERWRITE) { stg_stream(gx, 0.f); stg_stream(gy, 0.f); }
        return;
    }
    const int L = (int)x2, T = (int)y2;
    const int R = min(L + 1, f.W - 1), Bm = min(T + 1, f.H - 1);
    const float* cn = f.countp + b * f.count.b;
    const float* gox = f.goutp + b * f.out.b;
    const float* goy = gox + f.out.c;
    const int cx[4] = {L, R, L, R}, cy[4] = {T, T, Bm, Bm};
    float c[4], ax[4], ay[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c[k] = __ldg(cn + (int64_t)cy[k] * f.count.h + cx[k]);
        ax[k] = __ldg(gox + (int64_t)cy[k] * f.out.h + cx[k]);
        ay[k] = __ldg(goy + (int64_t)cy[k] * f.out.h + cx[k]);
    }
    float sx = OVERWRITE ? 0.f : *gx, sy = OVERWRITE ? 0.f : *gy;
#pragma unroll
    for (int k = 0; k < 4; ++k) sx += -ax[k] / c[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) sy += -ay[k] / c[k];
    *gx = sx;
    *gy = sy;
}

int wfp_splat_frame(cudaStream_t stream, const void* ctx, int b) {
    const WfpArgs& a = *static_cast<const WfpArgs*>(ctx);
    dim3 block(BX, BY, 1), grid((a.f.W + BX - 1) / BX, (a.f.H + BY - 1) / BY, 1);
    wfp_scatter_kernel<<<grid, block, 0, stream>>>(a, b);
    return 0;
}

int wfp_forward(cudaStream_t stream, const WfpArgs& a, int flags) {
    const FpArgs& f = a.f;
    if (f.B <= 0 || f.H <= 0 || f.W <= 0) return 0;
    if (f.B > 65535) return -1;
    DeviceGuard guard(f.flowp);
    if (!guard.ok) return -1;
    const bool ow = (flags & MEMC_B200_OVERWRITE) != 0, no_zero = (flags & MEMC_B200_NO_ZERO) != 0;
    if (ow && !no_zero && zero_fill(stream, a.wgtp, a.wgt, f.B, 1, f.H, f.W) != 0) return -1;
    dim3 block(BX, BY, 1), grid((f.W + BX - 1) / BX, (f.H + BY - 1) / BY, f.B);
    int r = 0;
    if (!(flags & MEMC_B200_NO_FAST)) {
        // FlowProjection's frame-by-frame driver: zero fills, this file's splat, average + occupancy masks, mask fill-hole
        // (the weight plane rides along in the average pass when it is dense like count: h-stride == W)
        const bool dense_w = a.wgt.h == f.W;
        r = fp_frames_fast(stream, f, ow, no_zero, false, wfp_splat_frame, &a, dense_w ? a.wgtp : nullptr, a.wgt.b);
        if (r < 0) return -1;
        if (r == 1 && dense_w) return 0;
    }
    if (r == 0) {
        if (ow && !no_zero) {
            if (zero_fill(stream, f.countp, f.count, f.B, 1, f.H, f.W) != 0) return -1;
            if (zero_fill(stream, f.outp, f.out, f.B, 2, f.H, f.W) != 0) return -1;
        }
        wfp_scatter_kernel<<<grid, block, 0, stream>>>(a, 0);
        count_launch();
        if (check_launch("WeightedFlowProjection scatter")) return -1;
        if (fp_average_fill(stream, f, 0, f.B, true) != 0) return -1;
    }
    wfp_weight_average_kernel<<<grid, block, 0, stream>>>(a);
    count_launch();
    return check_launch("WeightedFlowProjection weight average");
}

int wfp_backward(cudaStream_t stream, const WfpArgs& a, int flags) {
    const FpArgs& f = a.f;
    if (f.B <= 0 || f.H <= 0 || f.W <= 0) return 0;
    if (f.B > 65535) return -1;
    DeviceGuard guard(f.flowp);
    if (!guard.ok) return -1;
    dim3 block(BX, BY, 1), grid((f.W + BX - 1) / BX, (f.H + BY - 1) / BY, f.B);
    if (flags & MEMC_B200_OVERWRITE) wfp_bwd_kernel<true><<<grid, block, 0, stream>>>(a);
    else wfp_bwd_kernel<false><<<grid, block, 0, stream>>>(a);
    count_launch();
    return check_launch("WeightedFlowProjection backward");
}

}  // namespace

}  // namespace memc

using namespace memc;

extern "C" int memc_b200_weighted_flow_projection_forward(
    memc_stream_t stream, int batch, int h, int w, int fillhole, float threshold,
    memc_strides s_flow, memc_strides s_frame0, memc_strides s_frame1, memc_strides s_count, memc_strides s_weight,
    memc_strides s_out,
    const float* flow, const float* frame0, const float* frame1, float* count, float* weight, float* output, int flags) {
    WfpArgs a{};
    a.f.B = batch; a.f.H = h; a.f.W = w; a.f.fillhole = fillhole; a.threshold = threshold;
    a.f.flow = mk_view(s_flow); a.f.count = mk_view(s_count); a.f.out = mk_view(s_out);
    a.im0 = mk_view(s_frame0); a.im1 = mk_view(s_frame1); a.wgt = mk_view(s_weight);
    a.f.flowp = flow; a.im0p = frame0; a.im1p = frame1; a.f.countp = count; a.wgtp = weight; a.f.outp = output;
    return wfp_forward(stream, a, flags);
}

extern "C" int memc_b200_weighted_flow_projection_backward(
    memc_stream_t stream, int batch, int h, int w, float threshold,
    memc_strides s_flow, memc_strides s_frame0, memc_strides s_frame1, memc_strides s_count, memc_strides s_gout,
    memc_strides s_gi,
    const float* flow, const float* frame0, const float* frame1, const float* count, const float* gradoutput,
    float* gradinput, int flags) {
    WfpArgs a{};
    a.f.B = batch; a.f.H = h; a.f.W = w; a.threshold = threshold;
    a.f.flow = mk_view(s_flow); a.f.count = mk_view(s_count); a.f.out = mk_view(s_gout); a.f.gi = mk_view(s_gi);
    a.im0 = mk_view(s_frame0); a.im1 = mk_view(s_frame1);
    a.f.flowp = flow; a.im0p = frame0; a.im1p = frame1; a.f.countp = const_cast<float*>(count);
    a.f.goutp = gradoutput; a.f.gip = gradinput;
    return wfp_backward(stream, a, flags);
}

// Reference-named launchers (my_lib_kernel.h:222-255): output / gradoutput / gradinput1 use input1's strides; the frames'
// w-strides are honoured by the reference (my_lib_kernel.cu:2569-2572) and must be 1 here, like every other tensor's.
extern "C" int WeightedFlowProjection_gpu_forward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int channel, const int batch,
    const int fillhole, const float threshhold,
    const int i1b, const int i1c, const int i1h, const int i1w,
    const int i2b, const int i2c, const int i2h, const int i2w,
    const int i3b, const int i3c, const int i3h, const int i3w,
    const int cb, const int cc, const int ch, const int cw,
    const int wb, const int wc, const int wh, const int ww,
    const float* input1, const float* input2, const float* input3, float* count, float* weight, float* output) {
    (void)nElement; (void)cc; (void)wc;
    if (channel != 2 || i1w != 1 || i2w != 1 || i3w != 1 || cw != 1 || ww != 1) return -1;
    WfpArgs a{};
    a.f.B = batch; a.f.H = h; a.f.W = w; a.f.fillhole = fillhole; a.threshold = threshhold;
    a.f.flow = mk_view(i1b, i1c, i1h); a.f.count = mk_view(cb, 0, ch); a.f.out = a.f.flow;
    a.im0 = mk_view(i2b, i2c, i2h); a.im1 = mk_view(i3b, i3c, i3h); a.wgt = mk_view(wb, 0, wh);
    a.f.flowp = input1; a.im0p = input2; a.im1p = input3; a.f.countp = count; a.wgtp = weight; a.f.outp = output;
    return wfp_forward(stream, a, 0);
}

extern "C" int WeightedFlowProjection_gpu_backward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int channel, const int batch,
    const float threshhold,
    const int i1b, const int i1c, const int i1h, const int i1w,
    const int i2b, const int i2c, const int i2h, const int i2w,
    const int i3b, const int i3c, const int i3h, const int i3w,
    const int cb, const int cc, const int ch, const int cw,
    const int wb, const int wc, const int wh, const int ww,
    const float* input1, const float* input2, const float* input3, const float* count, const float* weight,
    const float* gradoutput, float* gradinput1) {
    (void)nElement; (void)cc; (void)wb; (void)wc; (void)wh; (void)weight;
    if (channel != 2 || i1w != 1 || i2w != 1 || i3w != 1 || cw != 1 || ww != 1) return -1;
    WfpArgs a{};
    a.f.B = batch; a.f.H = h; a.f.W = w; a.threshold = threshhold;
    a.f.flow = mk_view(i1b, i1c, i1h); a.f.count = mk_view(cb, 0, ch); a.f.out = a.f.flow; a.f.gi = a.f.flow;
    a.im0 = mk_view(i2b, i2c, i2h); a.im1 = mk_view(i3b, i3c, i3h);
    a.f.flowp = input1; a.im0p = input2; a.im1p = input3; a.f.countp = const_cast<float*>(count);
    a.f.goutp = gradoutput; a.f.gip = gradinput1;
    return wfp_backward(stream, a, 0);
}
