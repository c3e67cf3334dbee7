/*
 * memc_oracle.c -- CPU restatement of MEMC-Net's per-pixel motion-compensation ops.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path in
 * memc-net_b200/csrc/.  It may be imported, linked or executed only from tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The
 * product path never routes through it (and raises if its CUDA library is missing).
 *
 * PARITY PINNING: the reference ships no golden vectors for these ops,
 * so this restatement is pinned against the reference ITSELF: oracle/Makefile compiles
 * the reference's own my_package/src/my_lib.c unchanged (against th_stub/TH.h) into
 * oracle/_ref/libmemc_ref_cpu.so and tests/test_oracle_vs_reference.py requires the
 * float build of this file to agree with it BIT FOR BIT on seeded inputs; the committed
 * fixtures under tests/golden/ were generated from that same reference library
 * (tests/golden/make_golden.py).  Fill-hole has no CPU twin in the reference
 * (my_lib.c:1539-1543 prints "Not implemented"); it is restated from the CUDA source
 * my_lib_kernel.cu:1776-1833 and pinned on the GPU box against the reference kernels
 * recompiled for sm_100a (oracle/_ref/libmemc_ref_gpu.so).
 *
 * Two builds (oracle/Makefile): -DORACLE_REAL=float  -> liboracle_f32.so (operation order
 * of the reference CPU code, no FMA contraction) and -DORACLE_REAL=double ->
 * liboracle_f64.so (same geometry decisions in fp32, all products and sums in fp64; the
 * "true value" used to bound fp32 atomic-order noise).
 *
 * Layout: every tensor is dense NCHW fp32 (w-stride 1), the only layout the reference's
 * wrappers accept for these ops besides batch/channel-strided views
 * (my_lib_cuda.c:642-646).  Outputs are `real` (float or double per build).
 * Accumulating outputs (gi1/gi3 of FilterInterpolation, count/out of FlowProjection,
 * gi of FlowProjection backward, gi1 of Interpolation, all SeparableConv grads) are
 * ADDED into, exactly like the reference, so callers pass zero-filled buffers.
 *
 * All functions return 0 on success, -1 on a precondition violation (the reference's
 * error convention, my_lib_cuda.c:606-617).
 */
#include <math.h>
#include <stddef.h>

#ifndef ORACLE_REAL
#define ORACLE_REAL float
#endif
typedef ORACLE_REAL real;

#define ORACLE_API __attribute__((visibility("default")))

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int mini(int a, int b) { return a < b ? a : b; }

ORACLE_API int oracle_real_bytes(void) { return (int)sizeof(real); }

/* ------------------------------------------------------------------------------------
 * FilterInterpolation ("adaptive warp").
 * Follows my_lib.c:963-1073 (forward) / my_lib_kernel.cu:1121-1214.
 *
 * Geometry of one output pixel (h,w) with flow (fx,fy):
 *   x2 = w + fx, y2 = h + fy                                  (fp32)
 *   valid  <=>  0 <= x2 <= W-1  &&  0 <= y2 <= H-1  &&  |fx| < W/2  &&  |fy| < H/2
 *   ix = (int)x2, iy = (int)y2 (truncation; x2,y2 >= 0 so it is floor)
 *   window origin L = ix + 1 - fs/2, T = iy + 1 - fs/2, size fs x fs
 *   a tap (j,i) of the window reads image pixel (clamp(j,0,H-1), clamp(i,0,W-1)) and
 *   filter plane k = (j-T)*fs + (i-L) at the OUTPUT pixel (h,w)
 *   quadrant of a tap:  top <=> j <= iy,  left <=> i <= ix
 *   out = (1-a)(1-b) TL + a(1-b) TR + (1-a) b BL + a b BR,  a = x2-ix, b = y2-iy
 * invalid => out = in1[b,c,h,w]  (my_lib_kernel.cu:1209-1213)
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int valid;
    int ix, iy;      /* truncated target */
    int L, T;        /* window origin    */
    float alpha, beta;
} fi_geom;

static fi_geom fi_geometry(int h, int w, int H, int W, int fs, float fx, float fy)
{
    fi_geom g;
    float x2 = (float)w + fx;
    float y2 = (float)h + fy;
    g.valid = (x2 >= 0.0f && y2 >= 0.0f && x2 <= (float)(W - 1) && y2 <= (float)(H - 1) &&
               fabsf(fx) < (float)W / 2.0f && fabsf(fy) < (float)H / 2.0f);
    g.ix = g.iy = g.L = g.T = 0;
    g.alpha = g.beta = 0.0f;
    if (g.valid) {
        g.ix = (int)x2;
        g.iy = (int)y2;
        g.L = g.ix + 1 - fs / 2;
        g.T = g.iy + 1 - fs / 2;
        g.alpha = x2 - (float)g.ix;
        g.beta = y2 - (float)g.iy;
    }
    return g;
}

/* the four quadrant sums for one channel; quadrant order/tap order as my_lib.c:994-1040
 * (rows outer, columns inner inside each quadrant) */
static void fi_quadrants(const float *img /* [H,W] plane */, const float *filt /* batch base */,
                         size_t plane, size_t pix, int H, int W, int fs, const fi_geom *g,
                         real q[4])
{
    const int R = g->L + fs, Bm = g->T + fs;
    for (int qi = 0; qi < 4; ++qi) {
        const int top = (qi < 2), left = ((qi & 1) == 0);
        const int j0 = top ? g->T : g->iy + 1, j1 = top ? g->iy : Bm - 1;
        const int i0 = left ? g->L : g->ix + 1, i1 = left ? g->ix : R - 1;
        real acc = (real)0;
        for (int j = j0; j <= j1; ++j) {
            const int jj = clampi(j, 0, H - 1);
            for (int i = i0; i <= i1; ++i) {
                const int ii = clampi(i, 0, W - 1);
                const int k = (j - g->T) * fs + (i - g->L);
                acc += (real)img[(size_t)jj * W + ii] * (real)filt[(size_t)k * plane + pix];
            }
        }
        q[qi] = acc;
    }
}

ORACLE_API int oracle_filter_interpolation_forward(int B, int C, int H, int W, int fs,
                                                   const float *in1, const float *flow,
                                                   const float *filt, real *out)
{
    if (B < 0 || C < 0 || H <= 0 || W <= 0 || fs <= 0) return -1;
    const size_t plane = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
        const float *fb = filt + (size_t)b * fs * fs * plane;
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                const size_t pix = (size_t)h * W + w;
                const float fx = flow[((size_t)b * 2 + 0) * plane + pix];
                const float fy = flow[((size_t)b * 2 + 1) * plane + pix];
                const fi_geom g = fi_geometry(h, w, H, W, fs, fx, fy);
                for (int c = 0; c < C; ++c) {
                    const float *img = in1 + ((size_t)b * C + c) * plane;
                    real *o = out + ((size_t)b * C + c) * plane + pix;
                    if (!g.valid) { *o = (real)img[pix]; continue; }
                    real q[4];
                    fi_quadrants(img, fb, plane, pix, H, W, fs, &g, q);
                    const real a = (real)g.alpha, bt = (real)g.beta;
                    /* my_lib.c:1042-1046, left-to-right evaluation */
                    *o = ((real)1 - a) * ((real)1 - bt) * q[0] + a * ((real)1 - bt) * q[1] +
                         ((real)1 - a) * bt * q[2] + a * bt * q[3];
                }
            }
    }
    return 0;
}

/* Backward, my_lib.c:1151-1439 / my_lib_kernel.cu:1248-1515.
 * invalid pixel => contributes nothing at all (gi2 is not even written).
 * gi1 (+=, scatter to the clamped tap), gi3 (+=, own pixel), gi2 (=, own pixel). */
ORACLE_API int oracle_filter_interpolation_backward(int B, int C, int H, int W, int fs,
                                                    const float *in1, const float *flow,
                                                    const float *filt, const float *gout,
                                                    real *gi1, real *gi2, real *gi3)
{
    if (B < 0 || C < 0 || H <= 0 || W <= 0 || fs <= 0) return -1;
    const size_t plane = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
        const float *fb = filt + (size_t)b * fs * fs * plane;
        real *g3b = gi3 + (size_t)b * fs * fs * plane;
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                const size_t pix = (size_t)h * W + w;
                const float fx = flow[((size_t)b * 2 + 0) * plane + pix];
                const float fy = flow[((size_t)b * 2 + 1) * plane + pix];
                const fi_geom g = fi_geometry(h, w, H, W, fs, fx, fy);
                if (!g.valid) continue;
                const real a = (real)g.alpha, bt = (real)g.beta;
                const int R = g.L + fs, Bm = g.T + fs;

                /* steps 1+3: image and filter gradients (my_lib.c:1189-1253) */
                for (int c = 0; c < C; ++c) {
                    const float *img = in1 + ((size_t)b * C + c) * plane;
                    real *g1 = gi1 + ((size_t)b * C + c) * plane;
                    const real go = (real)gout[((size_t)b * C + c) * plane + pix];
                    for (int qi = 0; qi < 4; ++qi) {
                        const int top = (qi < 2), left = ((qi & 1) == 0);
                        /* go * {(1-a)|a} * {(1-b)|b}, evaluated left to right */
                        const real gq = go * (left ? (real)1 - a : a) * (top ? (real)1 - bt : bt);
                        const int j0 = top ? g.T : g.iy + 1, j1 = top ? g.iy : Bm - 1;
                        const int i0 = left ? g.L : g.ix + 1, i1 = left ? g.ix : R - 1;
                        for (int j = j0; j <= j1; ++j) {
                            const int jj = clampi(j, 0, H - 1);
                            for (int i = i0; i <= i1; ++i) {
                                const int ii = clampi(i, 0, W - 1);
                                const int k = (j - g.T) * fs + (i - g.L);
                                g1[(size_t)jj * W + ii] += gq * (real)fb[(size_t)k * plane + pix];
                                g3b[(size_t)k * plane + pix] += gq * (real)img[(size_t)jj * W + ii];
                            }
                        }
                    }
                }

                /* step 2: flow gradients (my_lib.c:1273-1414).  The reference writes
                 * gamma = 1 - beta and then uses (1 - gamma), NOT beta: keep that rounding. */
                real dx = (real)0, dy = (real)0;
                const real gam_y = (real)1 - bt, gam_x = (real)1 - a;
                for (int c = 0; c < C; ++c) {
                    const float *img = in1 + ((size_t)b * C + c) * plane;
                    const real go = (real)gout[((size_t)b * C + c) * plane + pix];
                    real q[4];
                    fi_quadrants(img, fb, plane, pix, H, W, fs, &g, q);
                    real t = (real)0;
                    t += gam_y * (q[1] - q[0]);
                    t += ((real)1 - gam_y) * (q[3] - q[2]);
                    dx += go * t;
                }
                for (int c = 0; c < C; ++c) {
                    const float *img = in1 + ((size_t)b * C + c) * plane;
                    const real go = (real)gout[((size_t)b * C + c) * plane + pix];
                    real q[4];
                    fi_quadrants(img, fb, plane, pix, H, W, fs, &g, q);
                    real t = (real)0;
                    t += gam_x * (q[2] - q[0]);
                    t += ((real)1 - gam_x) * (q[3] - q[1]);
                    dy += go * t;
                }
                gi2[((size_t)b * 2 + 0) * plane + pix] = dx;
                gi2[((size_t)b * 2 + 1) * plane + pix] = dy;
            }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------
 * FlowProjection: forward splat of -flow to the 4 integer neighbours of p + flow,
 * count, divide, optional hole fill.
 *   scatter + average : my_lib.c:1491-1536 (CUDA: my_lib_kernel.cu:1664-1690, 1730-1736)
 *   fill-hole         : my_lib_kernel.cu:1776-1833 ONLY (no CPU twin in the reference)
 * `count` is float (integer-valued, exact below 2^24) in both builds, as in the reference.
 * ---------------------------------------------------------------------------------- */
static inline int fp_valid(float x2, float y2, int H, int W)
{
    return x2 >= 0.0f && y2 >= 0.0f && x2 <= (float)(W - 1) && y2 <= (float)(H - 1);
}

ORACLE_API int oracle_flow_projection_forward(int B, int H, int W, const float *flow,
                                              float *count, real *out, int fillhole)
{
    if (B < 0 || H <= 0 || W <= 0) return -1;
    const size_t plane = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
        const float *fxp = flow + ((size_t)b * 2 + 0) * plane;
        const float *fyp = flow + ((size_t)b * 2 + 1) * plane;
        real *ox = out + ((size_t)b * 2 + 0) * plane;
        real *oy = out + ((size_t)b * 2 + 1) * plane;
        float *cnt = count + (size_t)b * plane;

        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                const float fx = fxp[(size_t)h * W + w], fy = fyp[(size_t)h * W + w];
                const float x2 = (float)w + fx, y2 = (float)h + fy;
                if (!fp_valid(x2, y2, H, W)) continue;
                const int L = (int)x2, T = (int)y2;
                const int R = mini(L + 1, W - 1), Bm = mini(T + 1, H - 1);
                const size_t cell[4] = {(size_t)T * W + L, (size_t)T * W + R,
                                        (size_t)Bm * W + L, (size_t)Bm * W + R};
                /* a clamped R/Bm makes the same cell appear twice: it is hit twice */
                for (int k = 0; k < 4; ++k) ox[cell[k]] += -(real)fx;
                for (int k = 0; k < 4; ++k) oy[cell[k]] += -(real)fy;
                for (int k = 0; k < 4; ++k) cnt[cell[k]] += 1.0f;
            }

        for (size_t p = 0; p < plane; ++p) {
            const float c = cnt[p];
            if (c > 0.0f) { ox[p] /= (real)c; oy[p] /= (real)c; }
        }

        if (!fillhole) continue;
        /* Holes (count <= 0) take the mean of the nearest non-hole to the left, right and
         * above.  The reference's downward search is `while(down_temp = 0.0f && ...)`
         * (my_lib_kernel.cu:1799): an assignment, so it never runs and "down" never
         * contributes.  Holes only read non-hole pixels, so the result is order-free. */
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                if (!(cnt[(size_t)h * W + w] <= 0.0f)) continue; /* `if(temp <= 0.0f)`: NaN is not a hole */
                int lo = w, ro = w, uo = h;
                float lt = 0.0f, rt = 0.0f, ut = 0.0f;
                while (lt == 0.0f && lo - 1 >= 0) { --lo; lt = cnt[(size_t)h * W + lo]; }
                while (rt == 0.0f && ro + 1 <= W - 1) { ++ro; rt = cnt[(size_t)h * W + ro]; }
                while (ut == 0.0f && uo - 1 >= 0) { --uo; ut = cnt[(size_t)uo * W + w]; }
                if (lt + rt + ut <= 0.0f) continue; /* nothing found: stays 0 */
                const real l = lt > 0.0f ? (real)1 : (real)0;
                const real r = rt > 0.0f ? (real)1 : (real)0;
                const real u = ut > 0.0f ? (real)1 : (real)0;
                const real den = l + r + u;
                real sx = (real)0, sy = (real)0;
                /* multiply-by-flag then add, in the reference's left/right/up order */
                if (lt > 0.0f) { sx += ox[(size_t)h * W + lo]; sy += oy[(size_t)h * W + lo]; }
                if (rt > 0.0f) { sx += ox[(size_t)h * W + ro]; sy += oy[(size_t)h * W + ro]; }
                if (ut > 0.0f) { sx += ox[(size_t)uo * W + w]; sy += oy[(size_t)uo * W + w]; }
                ox[(size_t)h * W + w] = sx / den;
                oy[(size_t)h * W + w] = sy / den;
            }
    }
    return 0;
}

/* my_lib.c:1590-1629: a gather over the same 4 cells, divided by the saved count;
 * ignores fill-hole.  gi is added into (reference uses +=). */
ORACLE_API int oracle_flow_projection_backward(int B, int H, int W, const float *flow,
                                               const float *count, const float *gout, real *gi)
{
    if (B < 0 || H <= 0 || W <= 0) return -1;
    const size_t plane = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
        const float *cnt = count + (size_t)b * plane;
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                const size_t pix = (size_t)h * W + w;
                const float fx = flow[((size_t)b * 2 + 0) * plane + pix];
                const float fy = flow[((size_t)b * 2 + 1) * plane + pix];
                const float x2 = (float)w + fx, y2 = (float)h + fy;
                if (!fp_valid(x2, y2, H, W)) continue;
                const int L = (int)x2, T = (int)y2;
                const int R = mini(L + 1, W - 1), Bm = mini(T + 1, H - 1);
                const size_t cell[4] = {(size_t)T * W + L, (size_t)T * W + R,
                                        (size_t)Bm * W + L, (size_t)Bm * W + R};
                for (int ch = 0; ch < 2; ++ch) {
                    const float *go = gout + ((size_t)b * 2 + ch) * plane;
                    real *g = gi + ((size_t)b * 2 + ch) * plane + pix;
                    for (int k = 0; k < 4; ++k) *g += -(real)go[cell[k]] / (real)cnt[cell[k]];
                }
            }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------
 * DepthFlowProjection: FlowProjection with a per-source weight w = input2[b,0,h,w]
 * (SURVEY.md section 8(f), rank 4).
 *   scatter + average : my_lib.c:1690-1735 (CUDA: my_lib_kernel.cu:2088-2118, 2158-2163)
 *   fill-hole         : my_lib_kernel.cu:2203-2259 ONLY (the CPU twin prints "Not implemented", my_lib.c:1738-1740);
 *                       the same walks as FlowProjection's, incl. the downward search that never runs (:2228)
 *   backward          : my_lib.c:1808-1872 (CUDA: my_lib_kernel.cu:2297-2357)
 * `count` (the accumulated weight) is float in both builds, as in the reference; in the float build every
 * expression keeps the reference's order of operations (-w * f, go * w / count, go / count * (f - out)).
 * ---------------------------------------------------------------------------------- */
ORACLE_API int oracle_depth_flow_projection_forward(int B, int H, int W, const float *flow, const float *depth,
                                                    float *count, real *out, int fillhole)
{
    if (B < 0 || H <= 0 || W <= 0) return -1;
    const size_t plane = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
        const float *fxp = flow + ((size_t)b * 2 + 0) * plane;
        const float *fyp = flow + ((size_t)b * 2 + 1) * plane;
        const float *wp = depth + (size_t)b * plane;
        real *ox = out + ((size_t)b * 2 + 0) * plane;
        real *oy = out + ((size_t)b * 2 + 1) * plane;
        float *cnt = count + (size_t)b * plane;

        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                const float fx = fxp[(size_t)h * W + w], fy = fyp[(size_t)h * W + w];
                const float x2 = (float)w + fx, y2 = (float)h + fy;
                if (!fp_valid(x2, y2, H, W)) continue;
                const int L = (int)x2, T = (int)y2;
                const int R = mini(L + 1, W - 1), Bm = mini(T + 1, H - 1);
                const size_t cell[4] = {(size_t)T * W + L, (size_t)T * W + R,
                                        (size_t)Bm * W + L, (size_t)Bm * W + R};
                const float wt = wp[(size_t)h * W + w];
                const real vx = -(real)wt * (real)fx, vy = -(real)wt * (real)fy;
                for (int k = 0; k < 4; ++k) ox[cell[k]] += vx;
                for (int k = 0; k < 4; ++k) oy[cell[k]] += vy;
                for (int k = 0; k < 4; ++k) cnt[cell[k]] += wt * 1.0f;
            }

        for (size_t p = 0; p < plane; ++p) {
            const float c = cnt[p];
            if (c > 0.0f) { ox[p] /= (real)c; oy[p] /= (real)c; }
        }

        if (!fillhole) continue;
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                if (!(cnt[(size_t)h * W + w] <= 0.0f)) continue; /* `if(temp <= 0.0f)`: NaN is not a hole */
                int lo = w, ro = w, uo = h;
                float lt = 0.0f, rt = 0.0f, ut = 0.0f;
                while (lt == 0.0f && lo - 1 >= 0) { --lo; lt = cnt[(size_t)h * W + lo]; }
                while (rt == 0.0f && ro + 1 <= W - 1) { ++ro; rt = cnt[(size_t)h * W + ro]; }
                while (ut == 0.0f && uo - 1 >= 0) { --uo; ut = cnt[(size_t)uo * W + w]; }
                if (lt + rt + ut <= 0.0f) continue;
                const real l = lt > 0.0f ? (real)1 : (real)0;
                const real r = rt > 0.0f ? (real)1 : (real)0;
                const real u = ut > 0.0f ? (real)1 : (real)0;
                const real den = l + r + u;
                real sx = (real)0, sy = (real)0;
                if (lt > 0.0f) { sx += ox[(size_t)h * W + lo]; sy += oy[(size_t)h * W + lo]; }
                if (rt > 0.0f) { sx += ox[(size_t)h * W + ro]; sy += oy[(size_t)h * W + ro]; }
                if (ut > 0.0f) { sx += ox[(size_t)uo * W + w]; sy += oy[(size_t)uo * W + w]; }
                ox[(size_t)h * W + w] = sx / den;
                oy[(size_t)h * W + w] = sy / den;
            }
    }
    return 0;
}

/* `fout` is the forward's output (float: what the forward stored).  gi1 / gi2 are added into. */
ORACLE_API int oracle_depth_flow_projection_backward(int B, int H, int W, const float *flow, const float *depth,
                                                     const float *count, const float *fout, const float *gout,
                                                     real *gi1, real *gi2)
{
    if (B < 0 || H <= 0 || W <= 0) return -1;
    const size_t plane = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
        const float *cnt = count + (size_t)b * plane;
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                const size_t pix = (size_t)h * W + w;
                const float f[2] = {flow[((size_t)b * 2 + 0) * plane + pix], flow[((size_t)b * 2 + 1) * plane + pix]};
                const float x2 = (float)w + f[0], y2 = (float)h + f[1];
                if (!fp_valid(x2, y2, H, W)) continue;
                const int L = (int)x2, T = (int)y2;
                const int R = mini(L + 1, W - 1), Bm = mini(T + 1, H - 1);
                const size_t cell[4] = {(size_t)T * W + L, (size_t)T * W + R,
                                        (size_t)Bm * W + L, (size_t)Bm * W + R};
                const float wt = depth[(size_t)b * plane + pix];
                for (int ch = 0; ch < 2; ++ch) {
                    const float *go = gout + ((size_t)b * 2 + ch) * plane;
                    real *g = gi1 + ((size_t)b * 2 + ch) * plane + pix;
                    for (int k = 0; k < 4; ++k) *g += -(real)go[cell[k]] * (real)wt / (real)cnt[cell[k]];
                }
                real *gw = gi2 + (size_t)b * plane + pix;
                for (int ch = 0; ch < 2; ++ch) {
                    const float *go = gout + ((size_t)b * 2 + ch) * plane;
                    const float *po = fout + ((size_t)b * 2 + ch) * plane;
                    for (int k = 0; k < 4; ++k)
                        *gw += -(real)go[cell[k]] / (real)cnt[cell[k]] * ((real)f[ch] - (real)po[cell[k]]);
                }
            }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------
 * WeightedFlowProjection: FlowProjection in which a source votes only if its brightness-constancy error is
 * <= threshold; the error is splatted and averaged into `weight` as well.
 *   scatter + average : my_lib.c:1941-2010 (CUDA: my_lib_kernel.cu:2557-2612, 2646-2654)
 *   fill-hole         : my_lib_kernel.cu:2705-2757 ONLY (CPU twin: "Not implemented"); output only, weight untouched
 *   backward          : my_lib.c:2106-2160 (CUDA: my_lib_kernel.cu:2800-2838)
 * The gate is an fp32 decision in BOTH builds, evaluated the way the reference's C code evaluates it: `fabs` is the
 * double function there, so each channel's term is (double)|d1 - d2| / 3.0 added to the float sum in double and rounded
 * back to float (my_lib.c:1960); the CUDA source has the same line but resolves fabs to the float overload -- the two
 * can differ in the last bit of the error, which only matters for a source sitting exactly on the threshold.
 * The error that is ACCUMULATED is that float value (both builds accumulate it in `real`).
 * ---------------------------------------------------------------------------------- */
static inline float wfp_error(int H, int W, const float *im0, const float *im1, size_t plane, int h, int w, float fx,
                              float fy)
{
    const float tx = (float)w + 2.0f * fx, ty = (float)h + 2.0f * fy;
    const float mx = tx < (float)W - 1.0f ? tx : (float)W - 1.0f, my = ty < (float)H - 1.0f ? ty : (float)H - 1.0f;
    const int x3 = (int)(mx > 0.0f ? mx : 0.0f), y3 = (int)(my > 0.0f ? my : 0.0f);
    float e = 0.0f;
    for (int c = 0; c < 3; ++c) {
        const float d1 = im0[c * plane + (size_t)h * W + w], d2 = im1[c * plane + (size_t)y3 * W + x3];
        e = (float)((double)e + fabs((double)(d1 - d2)) / 3.0);
    }
    e += 1e-8f;
    return e;
}

ORACLE_API int oracle_weighted_flow_projection_forward(int B, int H, int W, const float *flow, const float *im0,
                                                       const float *im1, float *count, real *weight, real *out,
                                                       int fillhole, float threshold)
{
    if (B < 0 || H <= 0 || W <= 0) return -1;
    const size_t plane = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
        const float *fxp = flow + ((size_t)b * 2 + 0) * plane;
        const float *fyp = flow + ((size_t)b * 2 + 1) * plane;
        real *ox = out + ((size_t)b * 2 + 0) * plane;
        real *oy = out + ((size_t)b * 2 + 1) * plane;
        real *wg = weight + (size_t)b * plane;
        float *cnt = count + (size_t)b * plane;

        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                const float fx = fxp[(size_t)h * W + w], fy = fyp[(size_t)h * W + w];
                const float x2 = (float)w + fx, y2 = (float)h + fy;
                if (!fp_valid(x2, y2, H, W)) continue;
                const float e = wfp_error(H, W, im0 + (size_t)b * 3 * plane, im1 + (size_t)b * 3 * plane, plane, h, w, fx, fy);
                if (!(e <= threshold)) continue;
                const int L = (int)x2, T = (int)y2;
                const int R = mini(L + 1, W - 1), Bm = mini(T + 1, H - 1);
                const size_t cell[4] = {(size_t)T * W + L, (size_t)T * W + R,
                                        (size_t)Bm * W + L, (size_t)Bm * W + R};
                for (int k = 0; k < 4; ++k) ox[cell[k]] += -(real)fx;
                for (int k = 0; k < 4; ++k) oy[cell[k]] += -(real)fy;
                for (int k = 0; k < 4; ++k) cnt[cell[k]] += 1.0f;
                for (int k = 0; k < 4; ++k) wg[cell[k]] += (real)e;
            }

        for (size_t p = 0; p < plane; ++p) {
            const float c = cnt[p];
            if (c > 0.0f) { ox[p] /= (real)c; oy[p] /= (real)c; wg[p] /= (real)c; }
        }

        if (!fillhole) continue;
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                if (!(cnt[(size_t)h * W + w] <= 0.0f)) continue;
                int lo = w, ro = w, uo = h;
                float lt = 0.0f, rt = 0.0f, ut = 0.0f;
                while (lt == 0.0f && lo - 1 >= 0) { --lo; lt = cnt[(size_t)h * W + lo]; }
                while (rt == 0.0f && ro + 1 <= W - 1) { ++ro; rt = cnt[(size_t)h * W + ro]; }
                while (ut == 0.0f && uo - 1 >= 0) { --uo; ut = cnt[(size_t)uo * W + w]; }
                if (lt + rt + ut <= 0.0f) continue;
                const real den = (real)((lt > 0.0f) + (rt > 0.0f) + (ut > 0.0f));
                real sx = (real)0, sy = (real)0;
                if (lt > 0.0f) { sx += ox[(size_t)h * W + lo]; sy += oy[(size_t)h * W + lo]; }
                if (rt > 0.0f) { sx += ox[(size_t)h * W + ro]; sy += oy[(size_t)h * W + ro]; }
                if (ut > 0.0f) { sx += ox[(size_t)uo * W + w]; sy += oy[(size_t)uo * W + w]; }
                ox[(size_t)h * W + w] = sx / den;
                oy[(size_t)h * W + w] = sy / den;
            }
    }
    return 0;
}

ORACLE_API int oracle_weighted_flow_projection_backward(int B, int H, int W, const float *flow, const float *im0,
                                                        const float *im1, const float *count, const float *gout,
                                                        real *gi, float threshold)
{
    if (B < 0 || H <= 0 || W <= 0) return -1;
    const size_t plane = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
        const float *cnt = count + (size_t)b * plane;
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                const size_t pix = (size_t)h * W + w;
                const float fx = flow[((size_t)b * 2 + 0) * plane + pix];
                const float fy = flow[((size_t)b * 2 + 1) * plane + pix];
                const float x2 = (float)w + fx, y2 = (float)h + fy;
                if (!fp_valid(x2, y2, H, W)) continue;
                const float e = wfp_error(H, W, im0 + (size_t)b * 3 * plane, im1 + (size_t)b * 3 * plane, plane, h, w, fx, fy);
                if (!(e <= threshold)) continue;
                const int L = (int)x2, T = (int)y2;
                const int R = mini(L + 1, W - 1), Bm = mini(T + 1, H - 1);
                const size_t cell[4] = {(size_t)T * W + L, (size_t)T * W + R,
                                        (size_t)Bm * W + L, (size_t)Bm * W + R};
                for (int ch = 0; ch < 2; ++ch) {
                    const float *go = gout + ((size_t)b * 2 + ch) * plane;
                    real *g = gi + ((size_t)b * 2 + ch) * plane + pix;
                    for (int k = 0; k < 4; ++k) *g += -(real)go[cell[k]] / (real)cnt[cell[k]];
                }
            }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------
 * PixelValue / PixelWeight / ReliableWeight: the 4x4 "pixel splat" family.
 * A source (h, w) lands at (w, h) + flow / 2; the 16 cells (T + m, L + n), m, n = -1..2, clamped to the frame, receive
 * the window weight g^2, g = 1 - ((beta - m)^2 + (alpha - n)^2) / (2 sigma_d^2), times flow_weight * input1[c] (mode 0,
 * PixelValue: my_lib.c:2680-2735), flow_weight (mode 1, PixelWeight: :2950-3003) or 1 (mode 2, ReliableWeight:
 * :3228-3284).  Backward (my_lib.c:2820-2884, 3085-3164, 3355-3397): sums over the same cells into the source pixel, in
 * the reference's loop nesting (m, n, c) and order of operations; modes 1 / 2 skip cells whose forward output is below
 * `threshold`.  Everything is added into the caller's buffers, like the reference.
 * ---------------------------------------------------------------------------------- */
static inline int px_geometry(int H, int W, const float *flow, size_t plane, int h, int w, int *L, int *T, float *alpha,
                              float *beta)
{
    const float fx = flow[(size_t)h * W + w], fy = flow[plane + (size_t)h * W + w];
    const float x2 = (float)w + fx / 2.0f, y2 = (float)h + fy / 2.0f;
    if (!(x2 >= 0.0f && y2 >= 0.0f && x2 <= (float)(W - 1) && y2 <= (float)(H - 1))) return 0;
    *L = (int)x2; *T = (int)y2;
    *alpha = x2 - (int)x2; *beta = y2 - (int)y2;
    return 1;
}

/* window weight before squaring, in `real` (float build: the reference's expression, operation for operation) */
static inline real px_window(float alpha, float beta, int m, int n, float sigma_d)
{
    const real bm = (real)beta - (real)m, an = (real)alpha - (real)n, sd = (real)sigma_d;
    return (real)1 - (bm * bm + an * an) / ((real)2 * sd * sd);
}

ORACLE_API int oracle_pixel_splat_forward(int mode, int B, int C, int H, int W, const float *in1, const float *flow,
                                          const float *fw, real *out, float sigma_d)
{
    if (B < 0 || H <= 0 || W <= 0 || mode < 0 || mode > 2) return -1;
    if (mode != 0) C = 1;
    const size_t plane = (size_t)H * W;
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                int L, T; float alpha, beta;
                if (!px_geometry(H, W, flow + (size_t)b * 2 * plane, plane, h, w, &L, &T, &alpha, &beta)) continue;
                for (int m = -1; m <= 2; ++m)
                    for (int n = -1; n <= 2; ++n) {
                        const int pm = clampi(m + T, 0, H - 1), pn = clampi(n + L, 0, W - 1);
                        real g = px_window(alpha, beta, m, n, sigma_d);
                        g = g * g;
                        if (mode == 2) { out[(size_t)b * plane + (size_t)pm * W + pn] += g; continue; }
                        const real f_w = (real)fw[(size_t)b * plane + (size_t)h * W + w];
                        if (mode == 1) { out[(size_t)b * plane + (size_t)pm * W + pn] += f_w * g; continue; }
                        for (int c = 0; c < C; ++c)
                            out[((size_t)b * C + c) * plane + (size_t)pm * W + pn] +=
                                f_w * g * (real)in1[((size_t)b * C + c) * plane + (size_t)h * W + w];
                    }
            }
    return 0;
}

/* fout: the forward's output (modes 1 / 2, threshold test); gi1 / gfw may be NULL where the mode has none */
ORACLE_API int oracle_pixel_splat_backward(int mode, int B, int C, int H, int W, const float *in1, const float *flow,
                                           const float *fw, const float *fout, const float *gout, real *gi1, real *gi3,
                                           real *gfw, float sigma_d, float threshold)
{
    if (B < 0 || H <= 0 || W <= 0 || mode < 0 || mode > 2) return -1;
    if (mode != 0) C = 1;
    const size_t plane = (size_t)H * W;
    const real s2 = (real)sigma_d * (real)sigma_d;
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                int L, T; float alpha, beta;
                if (!px_geometry(H, W, flow + (size_t)b * 2 * plane, plane, h, w, &L, &T, &alpha, &beta)) continue;
                const size_t pix = (size_t)h * W + w;
                real *gx = gi3 + (size_t)b * 2 * plane + pix, *gy = gx + plane;
                const real f_w = mode == 2 ? (real)1 : (real)fw[(size_t)b * plane + pix];
                for (int m = -1; m <= 2; ++m)
                    for (int n = -1; n <= 2; ++n) {
                        const int pm = clampi(m + T, 0, H - 1), pn = clampi(n + L, 0, W - 1);
                        const real g = px_window(alpha, beta, m, n, sigma_d);
                        const real dn = (real)n - (real)alpha, dm = (real)m - (real)beta;
                        if (mode == 0) {
                            for (int c = 0; c < C; ++c) {
                                const real go = (real)gout[((size_t)b * C + c) * plane + (size_t)pm * W + pn];
                                const real v = (real)in1[((size_t)b * C + c) * plane + pix];
                                gi1[((size_t)b * C + c) * plane + pix] += go * f_w * g * g;
                                gfw[(size_t)b * plane + pix] += go * g * g * v;
                                *gx += -go * f_w * v * g * dn / s2 * (real)2;
                                *gy += -go * f_w * v * g * dm / s2 * (real)2;
                            }
                            continue;
                        }
                        const real go = (real)gout[(size_t)b * plane + (size_t)pm * W + pn];
                        if (fout[(size_t)b * plane + (size_t)pm * W + pn] < threshold) continue;
                        if (mode == 1) {
                            gfw[(size_t)b * plane + pix] += go * g * g;
                            *gx += -go * f_w * g * dn / s2 * (real)2;
                            *gy += -go * f_w * g * dm / s2 * (real)2;
                        } else {   /* the reference's expression has no flow weight here: -go * g * (n - alpha) ... */
                            *gx += -go * g * dn / s2 * (real)2;
                            *gy += -go * g * dm / s2 * (real)2;
                        }
                    }
            }
    return 0;
}

/* ------------------------------------------------------------------------------------
 * WeightLayer: matching confidence (1 - err / lambda_e)^2, err = mean over the 3x3 neighbourhood and the channels of
 * |input1 around the pixel - input2 bilinearly sampled around pixel + flow|; 1e-4 where the target leaves the frame.
 *   forward  : my_lib.c:2297-2340 (CUDA: my_lib_kernel.cu:3048-3122)
 *   backward : my_lib.c:2468-2532 (CUDA: :3208-3322): the sign of every difference steers +-g into gi1 / gi2 / gi3
 * The float build follows the C source operation for operation (its `fabs` is the double function: the running sum is
 * rounded through double, my_lib.c:2325); the sign decisions are fp32 decisions in both builds.
 * ---------------------------------------------------------------------------------- */
static inline int wl_geometry(int H, int W, const float *flow, size_t plane, int h, int w, int *L, int *T, int *R, int *Bm,
                              float *alpha, float *beta)
{
    const float fx = flow[(size_t)h * W + w], fy = flow[plane + (size_t)h * W + w];
    const float x2 = (float)w + fx, y2 = (float)h + fy;
    if (!(x2 >= 0.0f && y2 >= 0.0f && x2 <= (float)(W - 1) && y2 <= (float)(H - 1))) return 0;
    *L = (int)x2; *T = (int)y2;
    *R = mini(*L + 1, W - 1); *Bm = mini(*T + 1, H - 1);
    *alpha = x2 - (int)x2; *beta = y2 - (int)y2;
    return 1;
}
/* the bilinear sample in fp32, as the reference evaluates it (drives the sign decisions in both builds) */
static inline float wl_target(float a, float b, float tl, float tr, float bl, float br)
{
    return (1 - a) * (1 - b) * tl + a * (1 - b) * tr + (1 - a) * b * bl + a * b * br;
}

ORACLE_API int oracle_weight_layer_forward(int B, int C, int H, int W, const float *in1, const float *in2, const float *flow,
                                           real *out, float lambda_e, float Nw)
{
    if (B < 0 || C <= 0 || H <= 0 || W <= 0) return -1;
    const size_t plane = (size_t)H * W;
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                int L, T, R, Bm; float a, bt;
                real *o = out + (size_t)b * plane + (size_t)h * W + w;
                if (!wl_geometry(H, W, flow + (size_t)b * 2 * plane, plane, h, w, &L, &T, &R, &Bm, &a, &bt)) { *o = (real)1e-4f; continue; }
                float err_f = 0.0f;
                real err = (real)0;
                for (int m = -1; m <= 1; ++m)
                    for (int n = -1; n <= 1; ++n) {
                        const int p1m = clampi(m + h, 0, H - 1), p1n = clampi(n + w, 0, W - 1);
                        const int mT = clampi(m + T, 0, H - 1), mB = clampi(m + Bm, 0, H - 1);
                        const int nL = clampi(n + L, 0, W - 1), nR = clampi(n + R, 0, W - 1);
                        for (int c = 0; c < C; ++c) {
                            const float *s = in2 + ((size_t)b * C + c) * plane;
                            const float tl = s[(size_t)mT * W + nL], tr = s[(size_t)mT * W + nR];
                            const float bl = s[(size_t)mB * W + nL], br = s[(size_t)mB * W + nR];
                            const float i_data = in1[((size_t)b * C + c) * plane + (size_t)p1m * W + p1n];
                            err_f = (float)((double)err_f + fabs((double)(i_data - wl_target(a, bt, tl, tr, bl, br))));
                            const real t64 = ((real)1 - a) * ((real)1 - bt) * tl + (real)a * ((real)1 - bt) * tr +
                                             ((real)1 - a) * (real)bt * bl + (real)a * (real)bt * br;
                            err += (real)fabs((double)((real)i_data - t64));
                        }
                    }
                if (sizeof(real) == sizeof(float)) {
                    err_f /= ((float)C * Nw * Nw);
                    *o = (real)((1 - err_f / lambda_e) * (1 - err_f / lambda_e));
                } else {
                    err /= ((real)C * Nw * Nw);
                    *o = ((real)1 - err / lambda_e) * ((real)1 - err / lambda_e);
                }
            }
    return 0;
}

/* fout = the forward's output; gi1 / gi2 / gi3 are added into */
ORACLE_API int oracle_weight_layer_backward(int B, int C, int H, int W, const float *in1, const float *in2, const float *flow,
                                            const float *fout, const float *gout, real *gi1, real *gi2, real *gi3,
                                            float lambda_e, float Nw)
{
    if (B < 0 || C <= 0 || H <= 0 || W <= 0) return -1;
    const size_t plane = (size_t)H * W;
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                int L, T, R, Bm; float af, btf;
                if (!wl_geometry(H, W, flow + (size_t)b * 2 * plane, plane, h, w, &L, &T, &R, &Bm, &af, &btf)) continue;
                const size_t pix = (size_t)h * W + w;
                const real a = (real)af, bt = (real)btf;
                const real go = (real)gout[(size_t)b * plane + pix];
                real ges;
                if (sizeof(real) == sizeof(float))
                    ges = (real)(-gout[(size_t)b * plane + pix] / (lambda_e * C * Nw * Nw) * 2 * sqrtf(fout[(size_t)b * plane + pix]));
                else
                    ges = -go / ((real)lambda_e * C * Nw * Nw) * 2 * (real)sqrt((double)fout[(size_t)b * plane + pix]);
                real *gx = gi3 + (size_t)b * 2 * plane + pix, *gy = gx + plane;
                for (int m = -1; m <= 1; ++m)
                    for (int n = -1; n <= 1; ++n) {
                        const int p1m = clampi(m + h, 0, H - 1), p1n = clampi(n + w, 0, W - 1);
                        const int mT = clampi(m + T, 0, H - 1), mB = clampi(m + Bm, 0, H - 1);
                        const int nL = clampi(n + L, 0, W - 1), nR = clampi(n + R, 0, W - 1);
                        for (int c = 0; c < C; ++c) {
                            const size_t ch = ((size_t)b * C + c) * plane;
                            const float *s = in2 + ch;
                            const float tl = s[(size_t)mT * W + nL], tr = s[(size_t)mT * W + nR];
                            const float bl = s[(size_t)mB * W + nL], br = s[(size_t)mB * W + nR];
                            const float i_data = in1[ch + (size_t)p1m * W + p1n];
                            const int above = i_data > wl_target(af, btf, tl, tr, bl, br);
                            const real s1 = above ? ges : -ges, s2 = above ? -ges : ges;
                            gi1[ch + (size_t)p1m * W + p1n] += s1;
                            gi2[ch + (size_t)mT * W + nL] += ((real)1 - a) * ((real)1 - bt) * s2;
                            gi2[ch + (size_t)mT * W + nR] += a * ((real)1 - bt) * s2;
                            gi2[ch + (size_t)mB * W + nL] += ((real)1 - a) * bt * s2;
                            gi2[ch + (size_t)mB * W + nR] += a * bt * s2;
                            real gamma = (real)1.0f - bt, t = (real)0;
                            t += gamma * ((real)tr - (real)tl);
                            t += ((real)1 - gamma) * ((real)br - (real)bl);
                            t = t * s2;
                            *gx += t;
                            gamma = (real)1.0f - a;
                            t = (real)0;
                            t += gamma * ((real)bl - (real)tl);
                            t += gamma * ((real)br - (real)tr);   /* gamma for both rows: my_lib.c:2524 / my_lib_kernel.cu:3310 */
                            t = t * s2;
                            *gy += t;
                        }
                    }
            }
    return 0;
}

/* ------------------------------------------------------------------------------------
 * SeparableConvFlow: the flow a pair of separable filters encodes (centroid of the taps minus (fs-1)/2) on the valid
 * region Ho x Wo = (H-fs+1) x (W-fs+1); -2000 where the taps sum to 0.  Follows the CUDA source my_lib_kernel.cu:52-80
 * (forward), :108-160 (backward); the CPU twin my_lib.c:53-66 divides by |sum| instead of the signed sum, so the two
 * reference implementations agree only for filters with a positive sum -- that is where this restatement is pinned
 * against the CPU twin.  gv is ASSIGNED, gh accumulated, as in the CUDA source (:131, :155); the CPU twin assigns both.
 * ---------------------------------------------------------------------------------- */
ORACLE_API int oracle_separable_conv_flow_forward(int B, int fs, int Ho, int Wo, const float *vert, const float *horiz,
                                                  real *flow)
{
    if (B < 0 || fs <= 0 || Ho <= 0 || Wo <= 0) return -1;
    const size_t plane = (size_t)Ho * Wo;
    const double half = ((double)(float)fs - 1.0) / 2.0;
    for (int b = 0; b < B; ++b)
        for (size_t p = 0; p < plane; ++p)
            for (int which = 0; which < 2; ++which) {   /* 0: vertical -> channel 1, 1: horizontal -> channel 0 */
                const float *f = (which ? horiz : vert) + (size_t)b * fs * plane + p;
                real cen = (real)0, sum = (real)0;
                for (int k = 0; k < fs; ++k) { cen += (real)k * (real)f[k * plane]; sum += (real)f[k * plane]; }
                real *o = flow + ((size_t)b * 2 + (which ? 0 : 1)) * plane + p;
                if (sizeof(real) == sizeof(float)) *o = fabs((double)sum) > 0.0 ? (real)(float)((double)(float)(cen / sum) - half) : (real)-2000;
                else *o = fabs((double)sum) > 0.0 ? (real)((double)(cen / sum) - half) : (real)-2000;
            }
    return 0;
}

ORACLE_API int oracle_separable_conv_flow_backward(int B, int fs, int Ho, int Wo, const float *vert, const float *horiz,
                                                   const float *gflow, real *gv, real *gh)
{
    if (B < 0 || fs <= 0 || Ho <= 0 || Wo <= 0) return -1;
    const size_t plane = (size_t)Ho * Wo;
    for (int b = 0; b < B; ++b)
        for (size_t p = 0; p < plane; ++p)
            for (int which = 0; which < 2; ++which) {
                const float *f = (which ? horiz : vert) + (size_t)b * fs * plane + p;
                real cen = (real)0, sum = (real)0;
                for (int k = 0; k < fs; ++k) { cen += (real)k * (real)f[k * plane]; sum += (real)f[k * plane]; }
                if (!(fabs((double)sum) > 0.0)) continue;
                const real g = (real)gflow[((size_t)b * 2 + (which ? 0 : 1)) * plane + p];
                const real off = cen / (sum * sum);
                real *o = (which ? gh : gv) + (size_t)b * fs * plane + p;
                for (int k = 0; k < fs; ++k) {
                    const real v = g * ((real)k / sum - off);
                    if (which) o[k * plane] += v; else o[k * plane] = v;
                }
            }
    return 0;
}

/* ------------------------------------------------------------------------------------
 * Interpolation (plain bilinear backward warp), any channel count (the reference's
 * InterpolationCh variant is the same code with the channel==3 check removed,
 * my_lib_cuda.c:490,519).  my_lib.c:480-527 (fwd), 590-660 (bwd).
 * Note the validity test is x2 < W (not <= W-1) and out-of-range pixels give 0.
 * ---------------------------------------------------------------------------------- */
ORACLE_API int oracle_interpolation_forward(int B, int C, int H, int W, const float *in1,
                                            const float *flow, real *out)
{
    if (B < 0 || C < 0 || H <= 0 || W <= 0) return -1;
    const size_t plane = (size_t)H * W;
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                const size_t pix = (size_t)h * W + w;
                const float fx = flow[((size_t)b * 2 + 0) * plane + pix];
                const float fy = flow[((size_t)b * 2 + 1) * plane + pix];
                const float x2 = (float)w + fx, y2 = (float)h + fy;
                const int valid = x2 >= 0.0f && y2 >= 0.0f && x2 < (float)W && y2 < (float)H;
                int L = 0, T = 0, R = 0, Bm = 0;
                real a = 0, bt = 0;
                if (valid) {
                    L = (int)x2; T = (int)y2;
                    R = mini(L + 1, W - 1); Bm = mini(T + 1, H - 1);
                    a = (real)(x2 - (float)L); bt = (real)(y2 - (float)T);
                }
                for (int c = 0; c < C; ++c) {
                    const float *img = in1 + ((size_t)b * C + c) * plane;
                    real *o = out + ((size_t)b * C + c) * plane + pix;
                    if (!valid) { *o = (real)0; continue; }
                    const real TL = img[(size_t)T * W + L], TR = img[(size_t)T * W + R];
                    const real BL = img[(size_t)Bm * W + L], BR = img[(size_t)Bm * W + R];
                    *o = ((real)1 - a) * ((real)1 - bt) * TL + a * ((real)1 - bt) * TR +
                         ((real)1 - a) * bt * BL + a * bt * BR;
                }
            }
    return 0;
}

ORACLE_API int oracle_interpolation_backward(int B, int C, int H, int W, const float *in1,
                                             const float *flow, const float *gout, real *gi1,
                                             real *gi2)
{
    if (B < 0 || C < 0 || H <= 0 || W <= 0) return -1;
    const size_t plane = (size_t)H * W;
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                const size_t pix = (size_t)h * W + w;
                const float fx = flow[((size_t)b * 2 + 0) * plane + pix];
                const float fy = flow[((size_t)b * 2 + 1) * plane + pix];
                const float x2 = (float)w + fx, y2 = (float)h + fy;
                if (!(x2 >= 0.0f && y2 >= 0.0f && x2 < (float)W && y2 < (float)H)) continue;
                const int L = (int)x2, T = (int)y2;
                const int R = mini(L + 1, W - 1), Bm = mini(T + 1, H - 1);
                const real a = (real)(x2 - (float)L), bt = (real)(y2 - (float)T);
                for (int c = 0; c < C; ++c) {
                    real *g1 = gi1 + ((size_t)b * C + c) * plane;
                    const real go = (real)gout[((size_t)b * C + c) * plane + pix];
                    g1[(size_t)T * W + L] += go * ((real)1 - a) * ((real)1 - bt);
                    g1[(size_t)T * W + R] += go * a * ((real)1 - bt);
                    g1[(size_t)Bm * W + L] += go * ((real)1 - a) * bt;
                    g1[(size_t)Bm * W + R] += go * a * bt;
                }
                /* gamma = Bm - y2 (my_lib.c:622): negative-capable when Bm was clamped */
                real gam = (real)((float)Bm - y2), dx = (real)0, dy = (real)0;
                for (int c = 0; c < C; ++c) {
                    const float *img = in1 + ((size_t)b * C + c) * plane;
                    const real go = (real)gout[((size_t)b * C + c) * plane + pix];
                    real t = (real)0;
                    t += gam * ((real)img[(size_t)T * W + R] - (real)img[(size_t)T * W + L]);
                    t += ((real)1 - gam) * ((real)img[(size_t)Bm * W + R] - (real)img[(size_t)Bm * W + L]);
                    dx += go * t;
                }
                gam = (real)((float)R - x2);
                for (int c = 0; c < C; ++c) {
                    const float *img = in1 + ((size_t)b * C + c) * plane;
                    const real go = (real)gout[((size_t)b * C + c) * plane + pix];
                    real t = (real)0;
                    t += gam * ((real)img[(size_t)Bm * W + L] - (real)img[(size_t)T * W + L]);
                    t += ((real)1 - gam) * ((real)img[(size_t)Bm * W + R] - (real)img[(size_t)T * W + R]);
                    dy += go * t;
                }
                gi2[((size_t)b * 2 + 0) * plane + pix] = dx;
                gi2[((size_t)b * 2 + 1) * plane + pix] = dy;
            }
    return 0;
}

/* ------------------------------------------------------------------------------------
 * SeparableConv: per-pixel separable fs x fs local convolution over the valid region.
 * my_lib.c:312-330 (fwd), 406-430 (bwd).  in1 [B,C,H,W]; vert/horiz [B,fs,Ho,Wo];
 * out [B,C,Ho,Wo] with Ho = H-fs+1, Wo = W-fs+1.
 * ---------------------------------------------------------------------------------- */
ORACLE_API int oracle_separable_conv_forward(int B, int C, int H, int W, int fs,
                                             const float *in1, const float *vert,
                                             const float *horiz, real *out)
{
    const int Ho = H - fs + 1, Wo = W - fs + 1;
    if (B < 0 || C < 0 || fs <= 0 || Ho <= 0 || Wo <= 0) return -1;
    const size_t ip = (size_t)H * W, op = (size_t)Ho * Wo;
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < Ho; ++h)
            for (int w = 0; w < Wo; ++w)
                for (int c = 0; c < C; ++c) {
                    real acc = (real)0;
                    for (int y = 0; y < fs; ++y)
                        for (int x = 0; x < fs; ++x) {
                            const real t1 = in1[((size_t)b * C + c) * ip + (size_t)(h + y) * W + (w + x)];
                            const real t2 = vert[((size_t)b * fs + y) * op + (size_t)h * Wo + w];
                            const real t3 = horiz[((size_t)b * fs + x) * op + (size_t)h * Wo + w];
                            acc += t1 * t2 * t3;
                        }
                    out[((size_t)b * C + c) * op + (size_t)h * Wo + w] = acc;
                }
    return 0;
}

ORACLE_API int oracle_separable_conv_backward(int B, int C, int H, int W, int fs,
                                              const float *in1, const float *vert,
                                              const float *horiz, const float *gout, real *gi1,
                                              real *gi2, real *gi3)
{
    const int Ho = H - fs + 1, Wo = W - fs + 1;
    if (B < 0 || C < 0 || fs <= 0 || Ho <= 0 || Wo <= 0) return -1;
    const size_t ip = (size_t)H * W, op = (size_t)Ho * Wo;
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < Ho; ++h)
            for (int w = 0; w < Wo; ++w)
                for (int c = 0; c < C; ++c) {
                    const real go = gout[((size_t)b * C + c) * op + (size_t)h * Wo + w];
                    for (int y = 0; y < fs; ++y)
                        for (int x = 0; x < fs; ++x) {
                            const size_t i1 = ((size_t)b * C + c) * ip + (size_t)(h + y) * W + (w + x);
                            const size_t i2 = ((size_t)b * fs + y) * op + (size_t)h * Wo + w;
                            const size_t i3 = ((size_t)b * fs + x) * op + (size_t)h * Wo + w;
                            const real t1 = in1[i1], t2 = vert[i2], t3 = horiz[i3];
                            gi1[i1] += go * t2 * t3;
                            gi2[i2] += go * t1 * t3;
                            gi3[i3] += go * t1 * t2;
                        }
                }
    return 0;
}
