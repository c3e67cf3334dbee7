// flow_projection.cu -- FlowProjection: forward-splat -flow to the mid time step, count,
// average, optional hole fill; and its (gather) backward.
//
// Semantics: reference my_package/src/my_lib_kernel.cu:1630-1694 (scatter), :1696-1739
// (average), :1742-1836 (fill-hole), :1837-1901 (backward), sequenced as in the launcher
// :1905-1992.  CPU twin my_lib.c:1447-1634 (no fill-hole there).
#include "flow_projection.cuh"

namespace memc {


constexpr int BX = 32, BY = 8;

__device__ __forceinline__ bool fp_valid(float x2, float y2, int W, int H) {
    return x2 >= 0.0f && y2 >= 0.0f && x2 <= (float)(W - 1) && y2 <= (float)(H - 1);
}

// ------------------------------------------------------------------------------ scatter
// One source pixel per thread: 4 cells x (out.x, out.y, count).  When R or Bm is clamped
// the same cell is hit twice, as in the reference (my_lib_kernel.cu:1673-1689).
__global__ void __launch_bounds__(BX* BY) fp_scatter_kernel(const FpArgs p) {
    const int w = blockIdx.x * BX + threadIdx.x;
    const int h = blockIdx.y * BY + threadIdx.y;
    const int b = blockIdx.z;
    if (w >= p.W || h >= p.H) return;
    const float* fl = p.flowp + b * p.flow.b + h * p.flow.h + w;
    const float fx = ldg_stream(fl);
    const float fy = ldg_stream(fl + p.flow.c);
    const float x2 = (float)w + fx, y2 = (float)h + fy;
    if (!fp_valid(x2, y2, p.W, p.H)) return;
    const int L = (int)x2, T = (int)y2;
    const int R = min(L + 1, p.W - 1), Bm = min(T + 1, p.H - 1);
    float* ox = p.outp + b * p.out.b;
    float* oy = ox + p.out.c;
    float* cn = p.countp + b * p.count.b;
    const int64_t oT = (int64_t)T * p.out.h, oB = (int64_t)Bm * p.out.h;
    const int64_t cT = (int64_t)T * p.count.h, cB = (int64_t)Bm * p.count.h;
    red_add(ox + oT + L, -fx); red_add(ox + oT + R, -fx);
    red_add(ox + oB + L, -fx); red_add(ox + oB + R, -fx);
    red_add(oy + oT + L, -fy); red_add(oy + oT + R, -fy);
    red_add(oy + oB + L, -fy); red_add(oy + oB + R, -fy);
    red_add(cn + cT + L, 1.0f); red_add(cn + cT + R, 1.0f);
    red_add(cn + cB + L, 1.0f); red_add(cn + cB + R, 1.0f);
}

// ------------------------------------------------------------------------------ average
__global__ void __launch_bounds__(BX* BY) fp_average_kernel(const FpArgs p) {
    const int w = blockIdx.x * BX + threadIdx.x;
    const int h = blockIdx.y * BY + threadIdx.y;
    const int b = blockIdx.z;
    if (w >= p.W || h >= p.H) return;
    const float c = p.countp[b * p.count.b + h * p.count.h + w];
    if (c > 0.0f) {
        float* ox = p.outp + b * p.out.b + h * p.out.h + w;
        ox[0] = ox[0] / c;
        ox[p.out.c] = ox[p.out.c] / c;
    }
}

// ---------------------------------------------------------------------------- fill-hole
// A hole (count <= 0) becomes the mean of the nearest counted pixels to its left, right and
// above.  The reference's downward search never executes (`while(down_temp = 0.0f && ...)`,
// my_lib_kernel.cu:1799), so "down" never contributes; reproduced by not searching down.
// Holes read only non-hole pixels and write only themselves: no ordering hazard.
__global__ void __launch_bounds__(BX* BY) fp_fillhole_kernel(const FpArgs p) {
    const int w = blockIdx.x * BX + threadIdx.x;
    const int h = blockIdx.y * BY + threadIdx.y;
    const int b = blockIdx.z;
    if (w >= p.W || h >= p.H) return;
    const float* cn = p.countp + b * p.count.b;
    if (!(cn[h * p.count.h + w] <= 0.0f)) return;  // my_lib_kernel.cu:1778 `if(temp <= 0.0f)`: a NaN count is not a hole
    int lo = w, ro = w, uo = h;
    float lt = 0.f, rt = 0.f, ut = 0.f;
    const float* crow = cn + h * p.count.h;
    while (lt == 0.0f && lo - 1 >= 0) { --lo; lt = crow[lo]; }
    while (rt == 0.0f && ro + 1 <= p.W - 1) { ++ro; rt = crow[ro]; }
    while (ut == 0.0f && uo - 1 >= 0) { --uo; ut = cn[uo * p.count.h + w]; }
    if (lt + rt + ut <= 0.0f) return;
    const float l = lt > 0.0f ? 1.f : 0.f, r = rt > 0.0f ? 1.f : 0.f, u = ut > 0.0f ? 1.f : 0.f;
    const float den = l + r + u;
    float* ox = p.outp + b * p.out.b;
    float* oy = ox + p.out.c;
    float sx = 0.f, sy = 0.f;
    if (lt > 0.0f) { sx += ox[h * p.out.h + lo]; sy += oy[h * p.out.h + lo]; }
    if (rt > 0.0f) { sx += ox[h * p.out.h + ro]; sy += oy[h * p.out.h + ro]; }
    if (ut > 0.0f) { sx += ox[uo * p.out.h + w]; sy += oy[uo * p.out.h + w]; }
    ox[h * p.out.h + w] = sx / den;
    oy[h * p.out.h + w] = sy / den;
}

// ----------------------------------------------------------------------------- backward
// Two pixels per thread (rows h and h + BY of a 32 x 16 block): the op is a dependent chain
// flow -> 12 gathers -> store, so the second pixel's loads double the bytes in flight per thread.
template <bool OVERWRITE>
__global__ void __launch_bounds__(BX* BY, 6) fp_bwd_kernel(const FpArgs p) {
    const int w = blockIdx.x * BX + threadIdx.x;
    const int h0 = blockIdx.y * (2 * BY) + threadIdx.y;
    const int b = blockIdx.z;
    if (w >= p.W) return;
    float fx[2], fy[2];
    bool in[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int h = h0 + k * BY;
        in[k] = h < p.H;
        const float* fl = p.flowp + b * p.flow.b + (int64_t)min(h, p.H - 1) * p.flow.h + w;
        fx[k] = ldg_stream(fl);
        fy[k] = ldg_stream(fl + p.flow.c);
    }
    const float* cn = p.countp + b * p.count.b;
    const float* gox = p.goutp + b * p.out.b;
    const float* goy = gox + p.out.c;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (!in[k]) continue;
        const int h = h0 + k * BY;
        const float x2 = (float)w + fx[k], y2 = (float)h + fy[k];
        float* gx = p.gip + b * p.gi.b + (int64_t)h * p.gi.h + w;
        float* gy = gx + p.gi.c;
        if (!fp_valid(x2, y2, p.W, p.H)) {
            if (OVERWRITE) { stg_stream(gx, 0.f); stg_stream(gy, 0.f); }
            continue;
        }
        const int L = (int)x2, T = (int)y2;
        const int R = min(L + 1, p.W - 1), Bm = min(T + 1, p.H - 1);
        const int64_t oT = (int64_t)T * p.out.h, oB = (int64_t)Bm * p.out.h;
        const int64_t cT = (int64_t)T * p.count.h, cB = (int64_t)Bm * p.count.h;
        const float c0 = __ldg(cn + cT + L), c1 = __ldg(cn + cT + R);
        const float c2 = __ldg(cn + cB + L), c3 = __ldg(cn + cB + R);
        const float ax0 = __ldg(gox + oT + L), ax1 = __ldg(gox + oT + R), ax2 = __ldg(gox + oB + L), ax3 = __ldg(gox + oB + R);
        const float ay0 = __ldg(goy + oT + L), ay1 = __ldg(goy + oT + R), ay2 = __ldg(goy + oB + L), ay3 = __ldg(goy + oB + R);
        // same order as my_lib_kernel.cu:1879-1896: ((( -a0/c0 ) - a1/c1) - a2/c2) - a3/c3
        float sx = OVERWRITE ? 0.f : *gx;
        float sy = OVERWRITE ? 0.f : *gy;
        sx += -ax0 / c0; sx += -ax1 / c1; sx += -ax2 / c2; sx += -ax3 / c3;
        sy += -ay0 / c0; sy += -ay1 / c1; sy += -ay2 / c2; sy += -ay3 / c3;
        *gx = sx;
        *gy = sy;
    }
}

// average (+ fill-hole) over frames [b0, b0 + nb) of `a` -- also used by the fast path
int fp_average_fill(cudaStream_t stream, const FpArgs& a, int b0, int nb, bool do_average) {
    FpArgs f = a;
    f.flowp = a.flowp + b0 * a.flow.b;
    f.countp = a.countp + b0 * a.count.b;
    f.outp = a.outp + b0 * a.out.b;
    dim3 block(BX, BY, 1), grid((a.W + BX - 1) / BX, (a.H + BY - 1) / BY, nb);
    if (do_average) {
        fp_average_kernel<<<grid, block, 0, stream>>>(f);
        count_launch();
        if (check_launch("FlowProjection average")) return -1;
    }
    if (a.fillhole) {
        fp_fillhole_kernel<<<grid, block, 0, stream>>>(f);
        count_launch();
        if (check_launch("FlowProjection fill-hole")) return -1;
    }
    return 0;
}

static int fp_forward(cudaStream_t stream, const FpArgs& a, int flags) {
    if (a.B <= 0 || a.H <= 0 || a.W <= 0) return 0;
    DeviceGuard guard(a.flowp);
    if (!guard.ok) return -1;
    const bool ow = (flags & MEMC_B200_OVERWRITE) != 0;
    if (!(flags & MEMC_B200_NO_FAST)) {
        const int r = fp_forward_fast(stream, a, ow, (flags & MEMC_B200_NO_ZERO) != 0, (flags >> 16) & 0xff);
        if (r != 0) return r < 0 ? -1 : 0;
    }
    if (ow && !(flags & MEMC_B200_NO_ZERO)) {
        if (zero_fill(stream, a.countp, a.count, a.B, 1, a.H, a.W) != 0) return -1;
        if (zero_fill(stream, a.outp, a.out, a.B, 2, a.H, a.W) != 0) return -1;
    }
    dim3 block(BX, BY, 1), grid((a.W + BX - 1) / BX, (a.H + BY - 1) / BY, a.B);
    fp_scatter_kernel<<<grid, block, 0, stream>>>(a);
    count_launch();
    if (check_launch("FlowProjection scatter")) return -1;
    fp_average_kernel<<<grid, block, 0, stream>>>(a);
    count_launch();
    if (check_launch("FlowProjection average")) return -1;
    if (a.fillhole) {
        fp_fillhole_kernel<<<grid, block, 0, stream>>>(a);
        count_launch();
        if (check_launch("FlowProjection fill-hole")) return -1;
    }
    return 0;
}

static int fp_backward(cudaStream_t stream, const FpArgs& a, int flags) {
    if (a.B <= 0 || a.H <= 0 || a.W <= 0) return 0;
    DeviceGuard guard(a.flowp);
    if (!guard.ok) return -1;
    dim3 block(BX, BY, 1), grid((a.W + BX - 1) / BX, (a.H + 2 * BY - 1) / (2 * BY), a.B);
    if (flags & MEMC_B200_OVERWRITE) fp_bwd_kernel<true><<<grid, block, 0, stream>>>(a);
    else fp_bwd_kernel<false><<<grid, block, 0, stream>>>(a);
    count_launch();
    return check_launch("FlowProjection backward");
}

}  // namespace memc

using namespace memc;

extern "C" int memc_b200_flow_projection_forward(
    memc_stream_t stream, int batch, int h, int w, int fillhole,
    memc_strides s_flow, memc_strides s_count, memc_strides s_out,
    const float* flow, float* count, float* output, int flags) {
    FpArgs a{};
    a.B = batch; a.H = h; a.W = w; a.fillhole = fillhole;
    a.flow = mk_view(s_flow); a.count = mk_view(s_count); a.out = mk_view(s_out);
    a.flowp = flow; a.countp = count; a.outp = output;
    return fp_forward(stream, a, flags);
}

extern "C" int memc_b200_flow_projection_backward(
    memc_stream_t stream, int batch, int h, int w,
    memc_strides s_flow, memc_strides s_count, memc_strides s_gout, memc_strides s_gi,
    const float* flow, const float* count, const float* gradoutput, float* gradinput, int flags) {
    FpArgs a{};
    a.B = batch; a.H = h; a.W = w;
    a.flow = mk_view(s_flow); a.count = mk_view(s_count); a.out = mk_view(s_gout); a.gi = mk_view(s_gi);
    a.flowp = flow; a.countp = const_cast<float*>(count); a.goutp = gradoutput; a.gip = gradinput;
    return fp_backward(stream, a, flags);
}

// Reference-named launchers (my_lib_kernel.h:161-187): output / gradoutput / gradinput use
// input1's strides (my_lib_kernel.cu:1676, 1879).
extern "C" int FlowProjection_gpu_forward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int channel,
    const int batch, const int fillhole,
    const int i1b, const int i1c, const int i1h, const int i1w,
    const int cb, const int cc, const int ch, const int cw,
    const float* input1, float* count, float* output) {
    (void)nElement; (void)cc;
    if (channel != 2 || i1w != 1 || cw != 1) return -1;
    FpArgs a{};
    a.B = batch; a.H = h; a.W = w; a.fillhole = fillhole;
    a.flow = mk_view(i1b, i1c, i1h); a.count = mk_view(cb, 0, ch); a.out = a.flow;
    a.flowp = input1; a.countp = count; a.outp = output;
    return fp_forward(stream, a, 0);
}

extern "C" int FlowProjection_gpu_backward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int channel,
    const int batch,
    const int i1b, const int i1c, const int i1h, const int i1w,
    const int cb, const int cc, const int ch, const int cw,
    const float* input1, const float* count, const float* gradoutput, float* gradinput1) {
    (void)nElement; (void)cc;
    if (channel != 2 || i1w != 1 || cw != 1) return -1;
    FpArgs a{};
    a.B = batch; a.H = h; a.W = w;
    a.flow = mk_view(i1b, i1c, i1h); a.count = mk_view(cb, 0, ch); a.out = a.flow; a.gi = a.flow;
    a.flowp = input1; a.countp = const_cast<float*>(count); a.goutp = gradoutput; a.gip = gradinput1;
    return fp_backward(stream, a, 0);
}
