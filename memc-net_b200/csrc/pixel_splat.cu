// pixel_splat.cu -- PixelValue / PixelWeight / ReliableWeight ("the splat family").
// Every source pixel (h, w) with flow f = input3[b,:,h,w] lands at the INTERMEDIATE position (x2, y2) = (w, h) + f / 2 and,
// if that lies inside the frame, spreads over the 4x4 window around it, cell (T + m, L + n), m, n = -1..2 (clamped to the
// frame), with the window weight  g^2,  g = 1 - ((beta - m)^2 + (alpha - n)^2) / (2 sigma_d^2),  (alpha, beta) = frac(x2, y2):
//     PixelValue      output[b,c,cell] += flow_weight * g^2 * input1[b,c,h,w]      (c = 0..C-1)
//     PixelWeight     output[b,0,cell] += flow_weight * g^2
//     ReliableWeight  output[b,0,cell] += g^2
// The backward of each is a gather over the same 16 cells into the source pixel (gradients for input1, the flow and the flow
// weight as the op has them); PixelWeight / ReliableWeight skip cells whose forward output is below `threshhold`.
//
// Semantics: reference my_package/src/my_lib_kernel.cu:3398-3472 / 3532-3621 (PixelValue fwd / bwd), :3689-3754 / 3813-3896
// (PixelWeight), :3967-4036 / 4095-4176 (ReliableWeight), launchers :3474-3530, 3623-3687, 3756-3811, 3898-3965, 4038-4093,
// 4178-4245; CPU twins my_lib.c:2615-3400; FFI names my_lib_cuda.h:163-203 (Prowindow must be 2, PixelValue wants C = 3:
// checked by the FFI layer as in my_lib_cuda.c).  No Python class or caller in the reference.  tao_r is unused there too.
//
// Forward: one source per thread, 16 x C fire-and-forget reductions (RED.E.ADD.F32); the 16 cells of neighbouring sources
// overlap in the L2 (measured B=4 x 1080p: 0.67 / 0.23 / 0.23 ms = the legacy kernels': the L2 reduction rate is the bound;
// a shared-memory box as in the FilterInterpolation backward would cut 48 reductions per pixel to ~3, not built for ops
// nobody calls).  Backward: no atomics at all -- the reference's atomicAdd targets are the thread's own pixel, so the
// sums stay in registers and leave with one store each.
#include "memc_common.cuh"

namespace memc {

namespace {

constexpr int BX = 32, BY = 8;
enum { PX_VALUE = 0, PX_WEIGHT = 1, PX_RELIABLE = 2 };

struct PxArgs {
    int B, C, H, W;
    View in1, flow, fw, out;       // out: forward output; backward: the forward's output (threshold test)
    View gout, gi1, gi3, gfw;
    const float* in1p;
    const float* flowp;
    const float* fwp;
    float* outp;                   // fwd: written (accumulated into); bwd: read
    const float* goutp;
    float* gi1p;
    float* gi3p;
    float* gfwp;
    float sigma_d, threshold;
};

struct PxGeo {
    bool valid;
    int L, T;
    float alpha, beta;
};
__device__ __forceinline__ PxGeo px_geometry(const PxArgs& p, int b, int h, int w) {
    const float* fl = p.flowp + b * p.flow.b + (int64_t)h * p.flow.h + w;
    const float fx = ldg_stream(fl), fy = ldg_stream(fl + p.flow.c);
    const float x2 = (float)w + fx / 2.0f, y2 = (float)h + fy / 2.0f;  // the intermediate position (:3420)
    PxGeo g;
    g.valid = x2 >= 0.0f && y2 >= 0.0f && x2 <= (float)(p.W - 1) && y2 <= (float)(p.H - 1);
    g.L = g.valid ? (int)x2 : 0;
    g.T = g.valid ? (int)y2 : 0;
    g.alpha = x2 - (float)g.L;
    g.beta = y2 - (float)g.T;
    return g;
}
__device__ __forceinline__ float px_window(float alpha, float beta, int m, int n, float sigma_d) {
    return 1.0f - ((beta - (float)m) * (beta - (float)m) + (alpha - (float)n) * (alpha - (float)n)) / (2.0f * sigma_d * sigma_d);
}

template <int MODE>
__global__ void __launch_bounds__(BX* BY) px_fwd_kernel(const PxArgs p) {
    const int w = blockIdx.x * BX + threadIdx.x, h = blockIdx.y * BY + threadIdx.y, b = blockIdx.z;
    if (w >= p.W || h >= p.H) return;
    const PxGeo g = px_geometry(p, b, h, w);
    if (!g.valid) return;
    const float f_w = MODE == PX_RELIABLE ? 1.0f : ldg_stream(p.fwp + b * p.fw.b + (int64_t)h * p.fw.h + w);
    constexpr int CMAX = 4;
    float v[CMAX];
    const int C = MODE == PX_VALUE ? p.C : 1;
    if (MODE == PX_VALUE) {
#pragma unroll
        for (int c = 0; c < CMAX; ++c) v[c] = c < C ? ldg_stream(p.in1p + b * p.in1.b + c * p.in1.c + (int64_t)h * p.in1.h + w) : 0.f;
    }
    float* ob = p.outp + b * p.out.b;
#pragma unroll
    for (int m = -1; m <= 2; ++m) {
        const int pm = min(max(0, m + g.T), p.H - 1);
#pragma unroll
        for (int n = -1; n <= 2; ++n) {
            const int pn = min(max(0, n + g.L), p.W - 1);
            float gd = px_window(g.alpha, g.beta, m, n, p.sigma_d);
            gd = gd * gd;
            float* cell = ob + (int64_t)pm * p.out.h + pn;
            if (MODE == PX_VALUE) {
                if (C <= CMAX) {
#pragma unroll
                    for (int c = 0; c < CMAX; ++c)
                        if (c < C) red_add(cell + c * p.out.c, f_w * gd * v[c]);
                } else {
                    for (int c = 0; c < C; ++c)
                        red_add(cell + c * p.out.c, f_w * gd * __ldg(p.in1p + b * p.in1.b + c * p.in1.c + (int64_t)h * p.in1.h + w));
                }
            } else if (MODE == PX_WEIGHT) {
                red_add(cell, f_w * gd);
            } else {
                red_add(cell, gd);
            }
        }
    }
}

// gradients of the source pixel: sums over its 16 cells (and C channels).  Per cell and channel one gather and four FMAs:
// with q = gradoutput * g the reference's four terms (my_lib_kernel.cu:3598-3610) are
//     gradinput1[c] += f_w * (q g),  gradflow_weights += (q g) v[c],  gradinput3.x += -f_w * 2 / sigma_d^2 * (q v[c] (n - alpha)),
//     gradinput3.y likewise with (m - beta)
// and the constant factors are applied once, after the sums (the reference divides by sigma_d^2 in every term: 96 divisions
// per pixel at C = 3; the results differ by rounding only, far inside the 1e-5 bar).
template <int MODE, bool OVERWRITE>
__global__ void __launch_bounds__(BX* BY) px_bwd_kernel(const PxArgs p) {
    const int w = blockIdx.x * BX + threadIdx.x, h = blockIdx.y * BY + threadIdx.y, b = blockIdx.z;
    if (w >= p.W || h >= p.H) return;
    const int64_t pix3 = b * p.gi3.b + (int64_t)h * p.gi3.h + w;
    const int64_t pixw = MODE == PX_RELIABLE ? 0 : b * p.gfw.b + (int64_t)h * p.gfw.h + w;
    const PxGeo g = px_geometry(p, b, h, w);
    const int C = MODE == PX_VALUE ? p.C : 1;
    if (!g.valid) {
        if (OVERWRITE) {
            p.gi3p[pix3] = 0.f;
            p.gi3p[pix3 + p.gi3.c] = 0.f;
            if (MODE != PX_RELIABLE) p.gfwp[pixw] = 0.f;
            if (MODE == PX_VALUE)
                for (int c = 0; c < C; ++c) p.gi1p[b * p.gi1.b + c * p.gi1.c + (int64_t)h * p.gi1.h + w] = 0.f;
        }
        return;
    }
    const float f_w = MODE == PX_RELIABLE ? 1.0f : ldg_stream(p.fwp + b * p.fw.b + (int64_t)h * p.fw.h + w);
    const float k3 = -f_w * 2.0f / (p.sigma_d * p.sigma_d);
    float gx = 0.f, gy = 0.f, gw = 0.f;
    const float* gob = p.goutp + b * p.gout.b;
    if (MODE == PX_VALUE) {
        constexpr int CMAX = 4;
        for (int c0 = 0; c0 < C; c0 += CMAX) {  // channels in groups of 4 (one pass for the reference's C = 3)
            float v[CMAX], a1[CMAX];
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                v[c] = c0 + c < C ? ldg_stream(p.in1p + b * p.in1.b + (c0 + c) * p.in1.c + (int64_t)h * p.in1.h + w) : 0.f;
                a1[c] = 0.f;
            }
#pragma unroll
            for (int m = -1; m <= 2; ++m) {
                const int pm = min(max(0, m + g.T), p.H - 1);
#pragma unroll
                for (int n = -1; n <= 2; ++n) {
                    const int pn = min(max(0, n + g.L), p.W - 1);
                    const float gd = px_window(g.alpha, g.beta, m, n, p.sigma_d);
                    const float dn = (float)n - g.alpha, dm = (float)m - g.beta;
                    const float* cell = gob + (c0 * p.gout.c) + (int64_t)pm * p.gout.h + pn;
#pragma unroll
                    for (int c = 0; c < CMAX; ++c) {
                        if (c0 + c >= C) continue;
                        const float q = __ldg(cell + c * p.gout.c) * gd;
                        a1[c] = fmaf(q, gd, a1[c]);
                        gw = fmaf(q * gd, v[c], gw);
                        gx = fmaf(q * v[c], dn, gx);
                        gy = fmaf(q * v[c], dm, gy);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                if (c0 + c >= C) continue;
                float* o = p.gi1p + b * p.gi1.b + (c0 + c) * p.gi1.c + (int64_t)h * p.gi1.h + w;
                *o = OVERWRITE ? f_w * a1[c] : *o + f_w * a1[c];
            }
        }
    } else {
        const float* fob = p.outp + b * p.out.b;
#pragma unroll
        for (int m = -1; m <= 2; ++m) {
            const int pm = min(max(0, m + g.T), p.H - 1);
#pragma unroll
            for (int n = -1; n <= 2; ++n) {
                const int pn = min(max(0, n + g.L), p.W - 1);
                const float gd = px_window(g.alpha, g.beta, m, n, p.sigma_d);
                const float go = __ldg(gob + (int64_t)pm * p.gout.h + pn);
                if (__ldg(fob + (int64_t)pm * p.out.h + pn) < p.threshold) continue;  // skip its gradients (:3863, :4145)
                const float q = go * gd;
                if (MODE == PX_WEIGHT) gw = fmaf(q, gd, gw);
                gx = fmaf(q, (float)n - g.alpha, gx);
                gy = fmaf(q, (float)m - g.beta, gy);
            }
        }
    }
    gx *= k3;
    gy *= k3;
    p.gi3p[pix3] = OVERWRITE ? gx : p.gi3p[pix3] + gx;
    p.gi3p[pix3 + p.gi3.c] = OVERWRITE ? gy : p.gi3p[pix3 + p.gi3.c] + gy;
    if (MODE != PX_RELIABLE) p.gfwp[pixw] = OVERWRITE ? gw : p.gfwp[pixw] + gw;
}

template <int MODE>
int px_forward(cudaStream_t stream, const PxArgs& a, int flags) {
    if (a.B <= 0 || a.H <= 0 || a.W <= 0 || (MODE == PX_VALUE && a.C <= 0)) return 0;
    if (a.B > 65535) return -1;
    DeviceGuard guard(a.flowp);
    if (!guard.ok) return -1;
    if ((flags & MEMC_B200_OVERWRITE) && !(flags & MEMC_B200_NO_ZERO) &&
        zero_fill(stream, a.outp, a.out, a.B, MODE == PX_VALUE ? a.C : 1, a.H, a.W) != 0)
        return -1;
    dim3 block(BX, BY, 1), grid((a.W + BX - 1) / BX, (a.H + BY - 1) / BY, a.B);
    px_fwd_kernel<MODE><<<grid, block, 0, stream>>>(a);
    count_launch();
    return check_launch("Pixel splat forward");
}

template <int MODE>
int px_backward(cudaStream_t stream, const PxArgs& a, int flags) {
    if (a.B <= 0 || a.H <= 0 || a.W <= 0 || (MODE == PX_VALUE && a.C <= 0)) return 0;
    if (a.B > 65535) return -1;
    DeviceGuard guard(a.flowp);
    if (!guard.ok) return -1;
    dim3 block(BX, BY, 1), grid((a.W + BX - 1) / BX, (a.H + BY - 1) / BY, a.B);
    if (flags & MEMC_B200_OVERWRITE) px_bwd_kernel<MODE, true><<<grid, block, 0, stream>>>(a);
    else px_bwd_kernel<MODE, false><<<grid, block, 0, stream>>>(a);
    count_launch();
    return check_launch("Pixel splat backward");
}

}  // namespace

}  // namespace memc

using namespace memc;

// ---- extended entry points (64-bit strides per tensor, flags) ----------------------------------------------------
extern "C" int memc_b200_pixel_value_forward(
    memc_stream_t stream, int batch, int channel, int h, int w, float sigma_d,
    memc_strides s_in1, memc_strides s_flow, memc_strides s_fw, memc_strides s_out,
    const float* input1, const float* flow, const float* flow_weights, float* output, int flags) {
    PxArgs a{};
    a.B = batch; a.C = channel; a.H = h; a.W = w; a.sigma_d = sigma_d;
    a.in1 = mk_view(s_in1); a.flow = mk_view(s_flow); a.fw = mk_view(s_fw); a.out = mk_view(s_out);
    a.in1p = input1; a.flowp = flow; a.fwp = flow_weights; a.outp = output;
    return px_forward<PX_VALUE>(stream, a, flags);
}

extern "C" int memc_b200_pixel_value_backward(
    memc_stream_t stream, int batch, int channel, int h, int w, float sigma_d,
    memc_strides s_in1, memc_strides s_flow, memc_strides s_fw, memc_strides s_gout, memc_strides s_gi1, memc_strides s_gi3,
    memc_strides s_gfw,
    const float* input1, const float* flow, const float* flow_weights, const float* gradoutput, float* gradinput1,
    float* gradinput3, float* gradflow_weights, int flags) {
    PxArgs a{};
    a.B = batch; a.C = channel; a.H = h; a.W = w; a.sigma_d = sigma_d;
    a.in1 = mk_view(s_in1); a.flow = mk_view(s_flow); a.fw = mk_view(s_fw); a.gout = mk_view(s_gout);
    a.gi1 = mk_view(s_gi1); a.gi3 = mk_view(s_gi3); a.gfw = mk_view(s_gfw);
    a.in1p = input1; a.flowp = flow; a.fwp = flow_weights; a.goutp = gradoutput;
    a.gi1p = gradinput1; a.gi3p = gradinput3; a.gfwp = gradflow_weights;
    return px_backward<PX_VALUE>(stream, a, flags);
}

extern "C" int memc_b200_pixel_weight_forward(
    memc_stream_t stream, int batch, int h, int w, float sigma_d,
    memc_strides s_flow, memc_strides s_fw, memc_strides s_out,
    const float* flow, const float* flow_weights, float* output, int flags) {
    PxArgs a{};
    a.B = batch; a.C = 1; a.H = h; a.W = w; a.sigma_d = sigma_d;
    a.flow = mk_view(s_flow); a.fw = mk_view(s_fw); a.out = mk_view(s_out);
    a.flowp = flow; a.fwp = flow_weights; a.outp = output;
    return px_forward<PX_WEIGHT>(stream, a, flags);
}

extern "C" int memc_b200_pixel_weight_backward(
    memc_stream_t stream, int batch, int h, int w, float sigma_d, float threshold,
    memc_strides s_flow, memc_strides s_fw, memc_strides s_out, memc_strides s_gout, memc_strides s_gi3, memc_strides s_gfw,
    const float* flow, const float* flow_weights, const float* output, const float* gradoutput, float* gradinput3,
    float* gradflow_weights, int flags) {
    PxArgs a{};
    a.B = batch; a.C = 1; a.H = h; a.W = w; a.sigma_d = sigma_d; a.threshold = threshold;
    a.flow = mk_view(s_flow); a.fw = mk_view(s_fw); a.out = mk_view(s_out); a.gout = mk_view(s_gout);
    a.gi3 = mk_view(s_gi3); a.gfw = mk_view(s_gfw);
    a.flowp = flow; a.fwp = flow_weights; a.outp = const_cast<float*>(output); a.goutp = gradoutput;
    a.gi3p = gradinput3; a.gfwp = gradflow_weights;
    return px_backward<PX_WEIGHT>(stream, a, flags);
}

extern "C" int memc_b200_reliable_weight_forward(
    memc_stream_t stream, int batch, int h, int w, float sigma_d,
    memc_strides s_flow, memc_strides s_out, const float* flow, float* output, int flags) {
    PxArgs a{};
    a.B = batch; a.C = 1; a.H = h; a.W = w; a.sigma_d = sigma_d;
    a.flow = mk_view(s_flow); a.out = mk_view(s_out);
    a.flowp = flow; a.outp = output;
    return px_forward<PX_RELIABLE>(stream, a, flags);
}

extern "C" int memc_b200_reliable_weight_backward(
    memc_stream_t stream, int batch, int h, int w, float sigma_d, float threshold,
    memc_strides s_flow, memc_strides s_out, memc_strides s_gout, memc_strides s_gi3,
    const float* flow, const float* output, const float* gradoutput, float* gradinput3, int flags) {
    PxArgs a{};
    a.B = batch; a.C = 1; a.H = h; a.W = w; a.sigma_d = sigma_d; a.threshold = threshold;
    a.flow = mk_view(s_flow); a.out = mk_view(s_out); a.gout = mk_view(s_gout); a.gi3 = mk_view(s_gi3);
    a.flowp = flow; a.outp = const_cast<float*>(output); a.goutp = gradoutput; a.gi3p = gradinput3;
    return px_backward<PX_RELIABLE>(stream, a, flags);
}

// ---- reference-named launchers (my_lib_kernel.h:297-398).  gradoutput / gradinput1 use input1's strides (PixelValue,
// my_lib_kernel.cu:3596-3599) resp. output's (PixelWeight / ReliableWeight, :3860, :4142); gradinput3 uses input3's and
// gradflow_weights flow_weights' strides; caller-zeroed buffers are accumulated into.
extern "C" int PixelValueLayer_gpu_forward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int channel, const int batch,
    const int i1b, const int i1c, const int i1h, const int i1w,
    const int i3b, const int i3c, const int i3h, const int i3w,
    const int fb, const int fc, const int fh, const int fww,
    const int ob, const int oc, const int oh, const int ow,
    const float* input1, const float* input3, const float* flow_weights, float* output,
    float sigma_d, float tao_r, float Prowindow) {
    (void)nElement; (void)fc; (void)tao_r; (void)Prowindow;
    if (i1w != 1 || i3w != 1 || fww != 1 || ow != 1) return -1;
    PxArgs a{};
    a.B = batch; a.C = channel; a.H = h; a.W = w; a.sigma_d = sigma_d;
    a.in1 = mk_view(i1b, i1c, i1h); a.flow = mk_view(i3b, i3c, i3h); a.fw = mk_view(fb, 0, fh); a.out = mk_view(ob, oc, oh);
    a.in1p = input1; a.flowp = input3; a.fwp = flow_weights; a.outp = output;
    return px_forward<PX_VALUE>(stream, a, 0);
}

extern "C" int PixelValueLayer_gpu_backward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int channel, const int batch,
    const int i1b, const int i1c, const int i1h, const int i1w,
    const int i3b, const int i3c, const int i3h, const int i3w,
    const int fb, const int fc, const int fh, const int fww,
    const int ob, const int oc, const int oh, const int ow,
    const float* input1, const float* input3, const float* flow_weights, const float* gradoutput, float* gradinput1,
    float* gradinput3, float* gradflow_weights, float sigma_d, float tao_r, float Prowindow) {
    (void)nElement; (void)fc; (void)ob; (void)oc; (void)oh; (void)ow; (void)tao_r; (void)Prowindow;
    if (i1w != 1 || i3w != 1 || fww != 1) return -1;
    PxArgs a{};
    a.B = batch; a.C = channel; a.H = h; a.W = w; a.sigma_d = sigma_d;
    a.in1 = mk_view(i1b, i1c, i1h); a.flow = mk_view(i3b, i3c, i3h); a.fw = mk_view(fb, 0, fh);
    a.gout = a.in1; a.gi1 = a.in1; a.gi3 = a.flow; a.gfw = a.fw;
    a.in1p = input1; a.flowp = input3; a.fwp = flow_weights; a.goutp = gradoutput;
    a.gi1p = gradinput1; a.gi3p = gradinput3; a.gfwp = gradflow_weights;
    return px_backward<PX_VALUE>(stream, a, 0);
}

extern "C" int PixelWeightLayer_gpu_forward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int batch,
    const int i3b, const int i3c, const int i3h, const int i3w,
    const int fb, const int fc, const int fh, const int fww,
    const int ob, const int oc, const int oh, const int ow,
    const float* input3, const float* flow_weights, float* output, float sigma_d, float tao_r, float Prowindow) {
    (void)nElement; (void)fc; (void)oc; (void)tao_r; (void)Prowindow;
    if (i3w != 1 || fww != 1 || ow != 1) return -1;
    PxArgs a{};
    a.B = batch; a.C = 1; a.H = h; a.W = w; a.sigma_d = sigma_d;
    a.flow = mk_view(i3b, i3c, i3h); a.fw = mk_view(fb, 0, fh); a.out = mk_view(ob, 0, oh);
    a.flowp = input3; a.fwp = flow_weights; a.outp = output;
    return px_forward<PX_WEIGHT>(stream, a, 0);
}

extern "C" int PixelWeightLayer_gpu_backward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int batch,
    const int i3b, const int i3c, const int i3h, const int i3w,
    const int fb, const int fc, const int fh, const int fww,
    const int ob, const int oc, const int oh, const int ow,
    const float* input3, const float* flow_weights, const float* output, const float* gradoutput, float* gradinput3,
    float* gradflow_weights, float threshhold, float sigma_d, float tao_r, float Prowindow) {
    (void)nElement; (void)fc; (void)oc; (void)tao_r; (void)Prowindow;
    if (i3w != 1 || fww != 1 || ow != 1) return -1;
    PxArgs a{};
    a.B = batch; a.C = 1; a.H = h; a.W = w; a.sigma_d = sigma_d; a.threshold = threshhold;
    a.flow = mk_view(i3b, i3c, i3h); a.fw = mk_view(fb, 0, fh); a.out = mk_view(ob, 0, oh);
    a.gout = a.out; a.gi3 = a.flow; a.gfw = a.fw;
    a.flowp = input3; a.fwp = flow_weights; a.outp = const_cast<float*>(output); a.goutp = gradoutput;
    a.gi3p = gradinput3; a.gfwp = gradflow_weights;
    return px_backward<PX_WEIGHT>(stream, a, 0);
}

extern "C" int ReliableWeightLayer_gpu_forward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int batch,
    const int i3b, const int i3c, const int i3h, const int i3w,
    const int ob, const int oc, const int oh, const int ow,
    const float* input3, float* output, float sigma_d, float tao_r, float Prowindow) {
    (void)nElement; (void)oc; (void)tao_r; (void)Prowindow;
    if (i3w != 1 || ow != 1) return -1;
    PxArgs a{};
    a.B = batch; a.C = 1; a.H = h; a.W = w; a.sigma_d = sigma_d;
    a.flow = mk_view(i3b, i3c, i3h); a.out = mk_view(ob, 0, oh);
    a.flowp = input3; a.outp = output;
    return px_forward<PX_RELIABLE>(stream, a, 0);
}

extern "C" int ReliableWeightLayer_gpu_backward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int batch,
    const int i3b, const int i3c, const int i3h, const int i3w,
    const int ob, const int oc, const int oh, const int ow,
    const float* input3, const float* output, const float* gradoutput, float* gradinput3,
    float threshhold, float sigma_d, float tao_r, float Prowindow) {
    (void)nElement; (void)oc; (void)tao_r; (void)Prowindow;
    if (i3w != 1 || ow != 1) return -1;
    PxArgs a{};
    a.B = batch; a.C = 1; a.H = h; a.W = w; a.sigma_d = sigma_d; a.threshold = threshhold;
    a.flow = mk_view(i3b, i3c, i3h); a.out = mk_view(ob, 0, oh); a.gout = a.out; a.gi3 = a.flow;
    a.flowp = input3; a.outp = const_cast<float*>(output); a.goutp = gradoutput; a.gi3p = gradinput3;
    return px_backward<PX_RELIABLE>(stream, a, 0);
}
