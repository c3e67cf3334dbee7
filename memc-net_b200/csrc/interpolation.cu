// interpolation.cu -- plain bilinear backward warp (Interpolation / InterpolationCh).
//
// Semantics: reference my_package/src/my_lib_kernel.cu:507-576 (forward), :578-680
// (backward); CPU twin my_lib.c:440-667.  The reference's "Ch" variant is the same kernel
// without the channel == 3 restriction of its wrapper (my_lib_cuda.c:373 vs :490), so one
// channel-generic implementation serves both names.
// Quirks kept: validity is x2 < W (not <= W-1); out-of-range pixels give 0 (and no
// gradient); the flow gradient uses gamma = Bm - y2 / R - x2 with the CLAMPED Bm / R.
#include "memc_common.cuh"

namespace memc {

struct IpArgs {
    int B, C, H, W;
    View in1, flow, out;  // out = output (fwd) / gradoutput (bwd)
    View gi1, gi2;
    const float* in1p;
    const float* flowp;
    float* outp;
    const float* goutp;
    float* gi1p;
    float* gi2p;
};

constexpr int BX = 32, BY = 8;

// 1 pixel / thread and a dependent chain (flow -> 4*C gathers -> store): latency bound, so the
// register cap buys occupancy (5 CTAs of 256 threads per SM) rather than unrolling depth.
template <int CT>  // CT = 3: RGB specialisation (fully unrolled); CT = 0: any channel count
__global__ void __launch_bounds__(BX* BY, 5) ip_fwd_kernel(const IpArgs p) {
    const int w = blockIdx.x * BX + threadIdx.x;
    const int h = blockIdx.y * BY + threadIdx.y;
    const int b = blockIdx.z;
    if (w >= p.W || h >= p.H) return;
    const float* fl = p.flowp + b * p.flow.b + h * p.flow.h + w;
    const float fx = ldg_stream(fl), fy = ldg_stream(fl + p.flow.c);
    const float x2 = (float)w + fx, y2 = (float)h + fy;
    float* ob = p.outp + b * p.out.b + h * p.out.h + w;
    const int C = CT ? CT : p.C;
    if (!(x2 >= 0.0f && y2 >= 0.0f && x2 < (float)p.W && y2 < (float)p.H)) {
        for (int c = 0; c < C; ++c) stg_stream(ob + c * p.out.c, 0.0f);
        return;
    }
    const int L = (int)x2, T = (int)y2;
    const int R = min(L + 1, p.W - 1), Bm = min(T + 1, p.H - 1);
    const float a = x2 - (float)L, bt = y2 - (float)T;
    const float wTL = (1.0f - a) * (1.0f - bt), wTR = a * (1.0f - bt);
    const float wBL = (1.0f - a) * bt, wBR = a * bt;
    const float* img = p.in1p + b * p.in1.b;
    const int64_t oT = (int64_t)T * p.in1.h, oB = (int64_t)Bm * p.in1.h;
    if (CT) {
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const float* im = img + c * p.in1.c;
            const float v = wTL * __ldg(im + oT + L) + wTR * __ldg(im + oT + R) + wBL * __ldg(im + oB + L) +
                            wBR * __ldg(im + oB + R);
            stg_stream(ob + c * p.out.c, v);
        }
    } else {
#pragma unroll 2
        for (int c = 0; c < C; ++c, img += p.in1.c) {
            const float v = wTL * __ldg(img + oT + L) + wTR * __ldg(img + oT + R) + wBL * __ldg(img + oB + L) +
                            wBR * __ldg(img + oB + R);
            stg_stream(ob + c * p.out.c, v);
        }
    }
}

template <bool OVERWRITE>
__global__ void __launch_bounds__(BX* BY, 4) ip_bwd_kernel(const IpArgs p) {
    const int w = blockIdx.x * BX + threadIdx.x;
    const int h = blockIdx.y * BY + threadIdx.y;
    const int b = blockIdx.z;
    if (w >= p.W || h >= p.H) return;
    const float* fl = p.flowp + b * p.flow.b + h * p.flow.h + w;
    const float fx = ldg_stream(fl), fy = ldg_stream(fl + p.flow.c);
    const float x2 = (float)w + fx, y2 = (float)h + fy;
    float* g2 = p.gi2p + b * p.gi2.b + h * p.gi2.h + w;
    if (!(x2 >= 0.0f && y2 >= 0.0f && x2 < (float)p.W && y2 < (float)p.H)) {
        if (OVERWRITE) { stg_stream(g2, 0.f); stg_stream(g2 + p.gi2.c, 0.f); }
        return;
    }
    const int L = (int)x2, T = (int)y2;
    const int R = min(L + 1, p.W - 1), Bm = min(T + 1, p.H - 1);
    const float a = x2 - (float)L, bt = y2 - (float)T;
    const float gam_y = (float)Bm - y2, gam_x = (float)R - x2;
    const float* img = p.in1p + b * p.in1.b;
    float* g1 = p.gi1p + b * p.gi1.b;
    const float* go = p.goutp + b * p.out.b + h * p.out.h + w;
    const int64_t oT = (int64_t)T * p.in1.h, oB = (int64_t)Bm * p.in1.h;
    const int64_t gT = (int64_t)T * p.gi1.h, gB = (int64_t)Bm * p.gi1.h;
    float dx = 0.f, dy = 0.f;
    for (int c = 0; c < p.C; ++c, img += p.in1.c, g1 += p.gi1.c) {
        const float gov = ldg_stream(go + c * p.out.c);
        red_add(g1 + gT + L, gov * (1.0f - a) * (1.0f - bt));
        red_add(g1 + gT + R, gov * a * (1.0f - bt));
        red_add(g1 + gB + L, gov * (1.0f - a) * bt);
        red_add(g1 + gB + R, gov * a * bt);
        const float TL = __ldg(img + oT + L), TR = __ldg(img + oT + R);
        const float BL = __ldg(img + oB + L), BR = __ldg(img + oB + R);
        dx = fmaf(gov, gam_y * (TR - TL) + (1.0f - gam_y) * (BR - BL), dx);
        dy = fmaf(gov, gam_x * (BL - TL) + (1.0f - gam_x) * (BR - TR), dy);
    }
    stg_stream(g2, dx);
    stg_stream(g2 + p.gi2.c, dy);
}

static int ip_forward(cudaStream_t stream, const IpArgs& a, int flags) {
    (void)flags;
    if (a.B <= 0 || a.C <= 0 || a.H <= 0 || a.W <= 0) return 0;
    DeviceGuard guard(a.in1p);
    if (!guard.ok) return -1;
    dim3 block(BX, BY, 1), grid((a.W + BX - 1) / BX, (a.H + BY - 1) / BY, a.B);
    if (a.C == 3) ip_fwd_kernel<3><<<grid, block, 0, stream>>>(a);
    else ip_fwd_kernel<0><<<grid, block, 0, stream>>>(a);
    count_launch();
    return check_launch("Interpolation forward");
}

static int ip_backward(cudaStream_t stream, const IpArgs& a, int flags) {
    if (a.B <= 0 || a.C <= 0 || a.H <= 0 || a.W <= 0) return 0;
    DeviceGuard guard(a.in1p);
    if (!guard.ok) return -1;
    const bool ow = (flags & MEMC_B200_OVERWRITE) != 0;
    if (ow && !(flags & MEMC_B200_NO_ZERO) && zero_fill(stream, a.gi1p, a.gi1, a.B, a.C, a.H, a.W) != 0) return -1;
    dim3 block(BX, BY, 1), grid((a.W + BX - 1) / BX, (a.H + BY - 1) / BY, a.B);
    if (ow) ip_bwd_kernel<true><<<grid, block, 0, stream>>>(a);
    else ip_bwd_kernel<false><<<grid, block, 0, stream>>>(a);
    count_launch();
    return check_launch("Interpolation backward");
}

}  // namespace memc

using namespace memc;

extern "C" int memc_b200_interpolation_forward(
    memc_stream_t stream, int batch, int channel, int h, int w,
    memc_strides s_in1, memc_strides s_flow, memc_strides s_out,
    const float* input1, const float* flow, float* output, int flags) {
    IpArgs a{};
    a.B = batch; a.C = channel; a.H = h; a.W = w;
    a.in1 = mk_view(s_in1); a.flow = mk_view(s_flow); a.out = mk_view(s_out);
    a.in1p = input1; a.flowp = flow; a.outp = output;
    return ip_forward(stream, a, flags);
}

extern "C" int memc_b200_interpolation_backward(
    memc_stream_t stream, int batch, int channel, int h, int w,
    memc_strides s_in1, memc_strides s_flow, memc_strides s_gout,
    memc_strides s_gi1, memc_strides s_gi2,
    const float* input1, const float* flow, const float* gradoutput,
    float* gradinput1, float* gradinput2, int flags) {
    IpArgs a{};
    a.B = batch; a.C = channel; a.H = h; a.W = w;
    a.in1 = mk_view(s_in1); a.flow = mk_view(s_flow); a.out = mk_view(s_gout);
    a.gi1 = mk_view(s_gi1); a.gi2 = mk_view(s_gi2);
    a.in1p = input1; a.flowp = flow; a.goutp = gradoutput; a.gi1p = gradinput1; a.gi2p = gradinput2;
    return ip_backward(stream, a, flags);
}

// Reference-named launchers (my_lib_kernel.h:67-132).  Note the reference launcher does NOT
// require w-stride 1 in its wrapper (my_lib_cuda.c:395-396) but its kernels index `+ w_i`
// unconditionally (my_lib_kernel.cu:543); a non-unit w-stride is rejected here.
static int ip_fwd_named(memc_stream_t stream, int w, int h, int channel, int batch,
                        int i1b, int i1c, int i1h, int i1w, int i2b, int i2c, int i2h, int i2w,
                        const float* input1, const float* input2, float* output) {
    if (i1w != 1 || i2w != 1) return -1;
    IpArgs a{};
    a.B = batch; a.C = channel; a.H = h; a.W = w;
    a.in1 = mk_view(i1b, i1c, i1h); a.flow = mk_view(i2b, i2c, i2h); a.out = a.in1;
    a.in1p = input1; a.flowp = input2; a.outp = output;
    return ip_forward(stream, a, 0);
}

static int ip_bwd_named(memc_stream_t stream, int w, int h, int channel, int batch,
                        int i1b, int i1c, int i1h, int i1w, int i2b, int i2c, int i2h, int i2w,
                        const float* input1, const float* input2, const float* gradoutput,
                        float* gradinput1, float* gradinput2) {
    if (i1w != 1 || i2w != 1) return -1;
    IpArgs a{};
    a.B = batch; a.C = channel; a.H = h; a.W = w;
    a.in1 = mk_view(i1b, i1c, i1h); a.flow = mk_view(i2b, i2c, i2h);
    a.out = a.in1; a.gi1 = a.in1; a.gi2 = a.flow;
    a.in1p = input1; a.flowp = input2; a.goutp = gradoutput; a.gi1p = gradinput1; a.gi2p = gradinput2;
    return ip_backward(stream, a, 0);
}

#define IP_STRIDE_PARAMS                                                                      \
    const int i1b, const int i1c, const int i1h, const int i1w, const int i2b, const int i2c, \
        const int i2h, const int i2w

extern "C" int InterpolationLayer_gpu_forward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int channel,
    const int batch, IP_STRIDE_PARAMS, const float* input1, const float* input2, float* output) {
    (void)nElement;
    return ip_fwd_named(stream, w, h, channel, batch, i1b, i1c, i1h, i1w, i2b, i2c, i2h, i2w, input1, input2, output);
}
extern "C" int InterpolationChLayer_gpu_forward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int channel,
    const int batch, IP_STRIDE_PARAMS, const float* input1, const float* input2, float* output) {
    (void)nElement;
    return ip_fwd_named(stream, w, h, channel, batch, i1b, i1c, i1h, i1w, i2b, i2c, i2h, i2w, input1, input2, output);
}
extern "C" int InterpolationLayer_gpu_backward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int channel,
    const int batch, IP_STRIDE_PARAMS, const float* input1, const float* input2,
    const float* gradoutput, float* gradinput1, float* gradinput2) {
    (void)nElement;
    return ip_bwd_named(stream, w, h, channel, batch, i1b, i1c, i1h, i1w, i2b, i2c, i2h, i2w, input1, input2,
                        gradoutput, gradinput1, gradinput2);
}
extern "C" int InterpolationChLayer_gpu_backward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int channel,
    const int batch, IP_STRIDE_PARAMS, const float* input1, const float* input2,
    const float* gradoutput, float* gradinput1, float* gradinput2) {
    (void)nElement;
    return ip_bwd_named(stream, w, h, channel, batch, i1b, i1c, i1h, i1w, i2b, i2c, i2h, i2w, input1, input2,
                        gradoutput, gradinput1, gradinput2);
}
