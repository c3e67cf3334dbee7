// weight_layer.cu -- WeightLayer: a per-pixel matching confidence.  For a source pixel (h, w) whose target (x2, y2) = (w, h) +
// flow lies inside the frame, the mean absolute difference over the 3x3 neighbourhood and the C channels between input1
// around (h, w) and input2 bilinearly sampled around (x2, y2) (same fractional position for all nine taps, integer corners
// shifted by the tap, everything clamped to the frame)
//     err = sum_{m,n,c} | input1[c, h+m, w+n] - bilinear(input2[c], corners + (m, n)) | / (C * Nw * Nw)
//     output = (1 - err / lambda_e)^2                 (1e-4 where the target is outside the frame)
// Backward: d output / d err = -2 sqrt(output) / lambda_e / (C Nw^2), pushed through |.| by the SIGN of each difference:
// +-g into gradinput1 around (h, w), -+g times the bilinear weights into gradinput2's four corners, and the spatial
// derivative of the bilinear sample into gradinput3 (with the reference's gamma quirk in the y component, :3308-3310).
//
// Semantics: reference my_package/src/my_lib_kernel.cu:3026-3125 (forward), :3189-3326 (backward), launchers :3127-3187,
// 3328-3396; CPU twin my_lib.c:2251-2614; FFI names my_lib_cuda.h:147-160 (Nw must be 3, C = 3: checked by the FFI layer).
// lambda_v is unused there too.  No Python class or caller in the reference.
//
// The sign decisions are taken on fp32 values whose last bit depends on how a compiler contracts the bilinear blend into
// FMAs: on inputs with near-ties two builds of the SAME source can disagree on a handful of signs (each flips one
// contribution of magnitude |g|).  tests/ compare the backward with robust statistics for that reason.
#include "memc_common.cuh"

namespace memc {

namespace {

constexpr int BX = 32, BY = 8;

struct WlArgs {
    int B, C, H, W;
    View in1, in2, flow, out;   // out: forward output [B,1,H,W] (bwd: read; gradoutput shares its strides)
    const float* in1p;
    const float* in2p;
    const float* flowp;
    float* outp;
    const float* goutp;
    float* gi1p;                // strides of in1
    float* gi2p;                // strides of in2
    float* gi3p;                // strides of flow
    float lambda_e, Nw;
};

struct WlGeo {
    bool valid;
    int L, T, R, Bm;
    float alpha, beta;
};
__device__ __forceinline__ WlGeo wl_geometry(const WlArgs& p, int b, int h, int w) {
    const float* fl = p.flowp + b * p.flow.b + (int64_t)h * p.flow.h + w;
    const float fx = ldg_stream(fl), fy = ldg_stream(fl + p.flow.c);
    const float x2 = (float)w + fx, y2 = (float)h + fy;
    WlGeo g;
    g.valid = x2 >= 0.0f && y2 >= 0.0f && x2 <= (float)(p.W - 1) && y2 <= (float)(p.H - 1);
    g.L = g.valid ? (int)x2 : 0;
    g.T = g.valid ? (int)y2 : 0;
    g.R = min(g.L + 1, p.W - 1);
    g.Bm = min(g.T + 1, p.H - 1);
    g.alpha = x2 - (float)g.L;
    g.beta = y2 - (float)g.T;
    return g;
}

// The nine taps' four bilinear corners overlap: with R = L + 1 and Bm = T + 1 they are the 4x4 cells (T-1..T+2) x (L-1..L+2)
// (clamped), 16 gathers per channel instead of 36; at the last column / row (R == L, Bm == T) the right / bottom corner IS
// the left / top one, which the dR / dB selects below reproduce (my_lib_kernel.cu:3066-3069: every index is clamped on its own).
struct WlPatch {
    float v[4][4];
};
__device__ __forceinline__ WlPatch wl_load_patch(const float* s, int64_t sh, const WlGeo& g, int H, int W) {
    WlPatch q;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int y = min(max(0, g.T - 1 + r), H - 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) q.v[r][k] = __ldg(s + (int64_t)y * sh + min(max(0, g.L - 1 + k), W - 1));
    }
    return q;
}

__global__ void __launch_bounds__(BX* BY) wl_fwd_kernel(const WlArgs p) {
    const int w = blockIdx.x * BX + threadIdx.x, h = blockIdx.y * BY + threadIdx.y, b = blockIdx.z;
    if (w >= p.W || h >= p.H) return;
    float* o = p.outp + b * p.out.b + (int64_t)h * p.out.h + w;
    const WlGeo g = wl_geometry(p, b, h, w);
    if (!g.valid) { *o = 1e-4f; return; }
    const float a = g.alpha, bt = g.beta;
    const bool dR = g.R != g.L, dB = g.Bm != g.T;
    const float* i1 = p.in1p + b * p.in1.b;
    const float* i2 = p.in2p + b * p.in2.b;
    float err = 0.0f;
    for (int c = 0; c < p.C; ++c) {   // (channel outermost: the sum is taken in a different order than the reference's)
        const WlPatch q = wl_load_patch(i2 + c * p.in2.c, p.in2.h, g, p.H, p.W);
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            const int p1m = min(max(0, m - 1 + h), p.H - 1);
#pragma unroll
            for (int n = 0; n < 3; ++n) {
                const int p1n = min(max(0, n - 1 + w), p.W - 1);
                const float tl = q.v[m][n], tr = dR ? q.v[m][n + 1] : q.v[m][n];
                const float bl = dB ? q.v[m + 1][n] : q.v[m][n], br = dB ? (dR ? q.v[m + 1][n + 1] : q.v[m + 1][n]) : tr;
                const float target = (1 - a) * (1 - bt) * tl + a * (1 - bt) * tr + (1 - a) * bt * bl + a * bt * br;
                err += fabsf(__ldg(i1 + c * p.in1.c + (int64_t)p1m * p.in1.h + p1n) - target);
            }
        }
    }
    err /= ((float)p.C * p.Nw * p.Nw);
    *o = (1 - err / p.lambda_e) * (1 - err / p.lambda_e);
}

// gradinput1 / gradinput2 are true scatters (a pixel's 3x3 neighbourhood and the four shifted corners overlap its
// neighbours'): fire-and-forget reductions, as in the reference; gradinput3 is the thread's own pixel (registers).
// (The forward's 4x4 register patch was measured here too: 125 registers, 2.92 ms against 2.43 ms -- the 135 reductions per
// pixel bound this kernel, not its gathers.)
template <bool OVERWRITE>
__global__ void __launch_bounds__(BX* BY) wl_bwd_kernel(const WlArgs p) {
    const int w = blockIdx.x * BX + threadIdx.x, h = blockIdx.y * BY + threadIdx.y, b = blockIdx.z;
    if (w >= p.W || h >= p.H) return;
    float* g3 = p.gi3p + b * p.flow.b + (int64_t)h * p.flow.h + w;
    const WlGeo g = wl_geometry(p, b, h, w);
    if (!g.valid) {
        if (OVERWRITE) { g3[0] = 0.f; g3[p.flow.c] = 0.f; }
        return;
    }
    const float a = g.alpha, bt = g.beta;
    const int64_t po = b * p.out.b + (int64_t)h * p.out.h + w;
    const float go = __ldg(p.goutp + po);
    const float ges = -go / (p.lambda_e * (float)p.C * p.Nw * p.Nw) * 2 * sqrtf(__ldg(p.outp + po));
    const float* i1 = p.in1p + b * p.in1.b;
    const float* i2 = p.in2p + b * p.in2.b;
    float* g1 = p.gi1p + b * p.in1.b;
    float* g2 = p.gi2p + b * p.in2.b;
    float gx = OVERWRITE ? 0.f : g3[0], gy = OVERWRITE ? 0.f : g3[p.flow.c];
    for (int m = -1; m <= 1; ++m) {
        const int p1m = min(max(0, m + h), p.H - 1);
        const int mT = min(max(0, m + g.T), p.H - 1), mB = min(max(0, m + g.Bm), p.H - 1);
        for (int n = -1; n <= 1; ++n) {
            const int p1n = min(max(0, n + w), p.W - 1);
            const int nL = min(max(0, n + g.L), p.W - 1), nR = min(max(0, n + g.R), p.W - 1);
            for (int c = 0; c < p.C; ++c) {
                const float* s = i2 + c * p.in2.c;
                const float tl = __ldg(s + (int64_t)mT * p.in2.h + nL), tr = __ldg(s + (int64_t)mT * p.in2.h + nR);
                const float bl = __ldg(s + (int64_t)mB * p.in2.h + nL), br = __ldg(s + (int64_t)mB * p.in2.h + nR);
                const float target = (1 - a) * (1 - bt) * tl + a * (1 - bt) * tr + (1 - a) * bt * bl + a * bt * br;
                const float i_data = __ldg(i1 + c * p.in1.c + (int64_t)p1m * p.in1.h + p1n);
                const bool above = i_data > target;
                const float s1 = above ? ges : -ges, s2 = above ? -ges : ges;
                red_add(g1 + c * p.in1.c + (int64_t)p1m * p.in1.h + p1n, s1);
                float* d = g2 + c * p.in2.c;
                red_add(d + (int64_t)mT * p.in2.h + nL, (1 - a) * (1 - bt) * s2);
                red_add(d + (int64_t)mT * p.in2.h + nR, a * (1 - bt) * s2);
                red_add(d + (int64_t)mB * p.in2.h + nL, (1 - a) * bt * s2);
                red_add(d + (int64_t)mB * p.in2.h + nR, a * bt * s2);
                float gamma = 1.0f - bt, t = 0.0f;
                t += gamma * (tr - tl);
                t += (1 - gamma) * (br - bl);
                gx += t * s2;
                gamma = 1.0f - a;
                t = 0.0f;
                t += gamma * (bl - tl);
                t += gamma * (br - tr);  // (the reference weights both rows with gamma here, my_lib_kernel.cu:3310)
                gy += t * s2;
            }
        }
    }
    g3[0] = gx;
    g3[p.flow.c] = gy;
}

int wl_launch(cudaStream_t stream, const WlArgs& a, bool backward, int flags) {
    if (a.B <= 0 || a.C <= 0 || a.H <= 0 || a.W <= 0) return 0;
    if (a.B > 65535) return -1;
    DeviceGuard guard(a.in1p);
    if (!guard.ok) return -1;
    dim3 block(BX, BY, 1), grid((a.W + BX - 1) / BX, (a.H + BY - 1) / BY, a.B);
    if (!backward) {
        wl_fwd_kernel<<<grid, block, 0, stream>>>(a);
    } else {
        if ((flags & MEMC_B200_OVERWRITE) && !(flags & MEMC_B200_NO_ZERO)) {
            if (zero_fill(stream, a.gi1p, a.in1, a.B, a.C, a.H, a.W) != 0) return -1;
            if (zero_fill(stream, a.gi2p, a.in2, a.B, a.C, a.H, a.W) != 0) return -1;
        }
        if (flags & MEMC_B200_OVERWRITE) wl_bwd_kernel<true><<<grid, block, 0, stream>>>(a);
        else wl_bwd_kernel<false><<<grid, block, 0, stream>>>(a);
    }
    count_launch();
    return check_launch(backward ? "WeightLayer backward" : "WeightLayer forward");
}

WlArgs wl_named(int w, int h, int channel, int batch, int i1b, int i1c, int i1h, int i2b, int i2c, int i2h, int i3b, int i3c, int i3h,
                int ob, int oh, float lambda_e, float Nw) {
    WlArgs a{};
    a.B = batch; a.C = channel; a.H = h; a.W = w; a.lambda_e = lambda_e; a.Nw = Nw;
    a.in1 = mk_view(i1b, i1c, i1h); a.in2 = mk_view(i2b, i2c, i2h); a.flow = mk_view(i3b, i3c, i3h); a.out = mk_view(ob, 0, oh);
    return a;
}

}  // namespace

}  // namespace memc

using namespace memc;

// Extended entry points: gradients share their inputs' strides, gradoutput the output's.
extern "C" int memc_b200_weight_layer_forward(
    memc_stream_t stream, int batch, int channel, int h, int w, float lambda_e, float Nw,
    memc_strides s_in1, memc_strides s_in2, memc_strides s_flow, memc_strides s_out,
    const float* input1, const float* input2, const float* flow, float* output, int flags) {
    WlArgs a{};
    a.B = batch; a.C = channel; a.H = h; a.W = w; a.lambda_e = lambda_e; a.Nw = Nw;
    a.in1 = mk_view(s_in1); a.in2 = mk_view(s_in2); a.flow = mk_view(s_flow); a.out = mk_view(s_out);
    a.in1p = input1; a.in2p = input2; a.flowp = flow; a.outp = output;
    return wl_launch(stream, a, false, flags);
}

extern "C" int memc_b200_weight_layer_backward(
    memc_stream_t stream, int batch, int channel, int h, int w, float lambda_e, float Nw,
    memc_strides s_in1, memc_strides s_in2, memc_strides s_flow, memc_strides s_out,
    const float* input1, const float* input2, const float* flow, const float* output, const float* gradoutput,
    float* gradinput1, float* gradinput2, float* gradinput3, int flags) {
    WlArgs a{};
    a.B = batch; a.C = channel; a.H = h; a.W = w; a.lambda_e = lambda_e; a.Nw = Nw;
    a.in1 = mk_view(s_in1); a.in2 = mk_view(s_in2); a.flow = mk_view(s_flow); a.out = mk_view(s_out);
    a.in1p = input1; a.in2p = input2; a.flowp = flow; a.outp = const_cast<float*>(output); a.goutp = gradoutput;
    a.gi1p = gradinput1; a.gi2p = gradinput2; a.gi3p = gradinput3;
    return wl_launch(stream, a, true, flags);
}

// Reference-named launchers (my_lib_kernel.h:257-294); caller-zeroed gradients are accumulated into.
extern "C" int WeightLayer_gpu_forward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int channel, const int batch,
    const int i1b, const int i1c, const int i1h, const int i1w,
    const int i2b, const int i2c, const int i2h, const int i2w,
    const int i3b, const int i3c, const int i3h, const int i3w,
    const int ob, const int oc, const int oh, const int ow,
    const float* input1, const float* input2, const float* input3, float* output,
    float lambda_e, float lambda_v, float Nw) {
    (void)nElement; (void)oc; (void)lambda_v;
    if (i1w != 1 || i2w != 1 || i3w != 1 || ow != 1) return -1;
    WlArgs a = wl_named(w, h, channel, batch, i1b, i1c, i1h, i2b, i2c, i2h, i3b, i3c, i3h, ob, oh, lambda_e, Nw);
    a.in1p = input1; a.in2p = input2; a.flowp = input3; a.outp = output;
    return wl_launch(stream, a, false, 0);
}

extern "C" int WeightLayer_gpu_backward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int channel, const int batch,
    const int i1b, const int i1c, const int i1h, const int i1w,
    const int i2b, const int i2c, const int i2h, const int i2w,
    const int i3b, const int i3c, const int i3h, const int i3w,
    const int ob, const int oc, const int oh, const int ow,
    const float* input1, const float* input2, const float* input3, const float* output, const float* gradoutput,
    float* gradinput1, float* gradinput2, float* gradinput3, float lambda_e, float lambda_v, float Nw) {
    (void)nElement; (void)oc; (void)lambda_v;
    if (i1w != 1 || i2w != 1 || i3w != 1 || ow != 1) return -1;
    WlArgs a = wl_named(w, h, channel, batch, i1b, i1c, i1h, i2b, i2c, i2h, i3b, i3c, i3h, ob, oh, lambda_e, Nw);
    a.in1p = input1; a.in2p = input2; a.flowp = input3; a.outp = const_cast<float*>(output); a.goutp = gradoutput;
    a.gi1p = gradinput1; a.gi2p = gradinput2; a.gi3p = gradinput3;
    return wl_launch(stream, a, true, 0);
}
