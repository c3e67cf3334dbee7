// filter_interpolation_bwd_rows.cu -- FilterInterpolation backward for sm_100a, fs = 4, C <= 4:
// the "tap-row lanes" kernel (round 2; replaces the one-pixel-per-thread fi_bwd_tma_kernel as the
// production backward; reference semantics my_lib_kernel.cu:1220-1518).
//
// What bounded the round-1 kernel (profiles/r01_ncu_bench_fi_bwd.txt): the L1 / shared-memory data pipe,
// 432 wavefronts per 32 pixels, of which 200 were bank-conflict replays -- 121 of them on the int32
// shared atomics (ATOMS.ADD), because with one pixel per lane a warp's 32 windows spread over ~44
// columns x ~10 rows of the box and equal banks (or equal addresses, which atomics cannot broadcast)
// collide ~3.6 ways.  The remedy is a different lane map, not a different data path:
//
//   lane = (pixel p, tap ROW j):  a warp instruction covers 8 neighbouring pixels of one image row, and
//   the 4 lanes of a pixel own the 4 rows of its 4x4 window.  Row pitch 72 words (= 8 mod 32) puts the
//   4 rows of a pixel 8 banks apart and the 8 pixels' windows next to each other: a step touches
//   8 x 4 words that are a permutation of the 32 banks whenever the 8 integer targets are distinct mod 8.
//   Modelled on the benchmark field (tools/bank_model.py --roles): 28.5 load / 36.2 atomic wavefronts per
//   32 pixels and channel, against 38.3 / 52.6 for one pixel per lane (measured: 44 / 58).
//
// The 4 lanes of a pixel need 4 different filter planes (4j+i) of the SAME pixel: with a dense
// [plane][y][x] tile those are the same bank.  The filter tile is therefore fetched through a rank-5
// tensor map that splits the plane index into (j, i) and puts j right after x (tma::make_map_taps): a
// strip of 8 pixels lands as [y][i][j][x], and a (pixel, tap row) warp reads 32 consecutive words.
// gradinput3 has the same shape, is staged IN PLACE over the filter strip (each lane owns the words it
// read) and leaves through one TMA store per strip; gradinput2 is reduced over the 4 lanes of a pixel with
// two shuffles, staged over the flow tile and stored by TMA as well.
//
// Other changes against round 1:
//   * the image / accumulation boxes are made of 4-row slabs and only the slabs the tile's windows
//     touch are loaded, zeroed, converted and flushed (TMA reduce-add per slab): a third less box traffic;
//   * filter taps and gradoutput are read from shared memory ONCE, into the registers of the lane that
//     uses them; the fixed-point bound M (max over pixels of max|gradoutput| * max|filter tap|) comes from those
//     registers (round 1 re-read the 16 + C planes for it);
//   * the warp geometry (validity, integer target, alpha / beta, fast-path test) is evaluated once per pixel in
//     the row-segment view (lane = x) and handed to the pixel's 4 tap-row lanes with 3 shuffles: with 4 lanes per
//     pixel every per-pixel instruction costs 4 issue slots, and geometry was a third of the first version;
// Fixed-point accumulation of gradinput1: as in round 1 (per-tile power-of-two scale, K = TW*TH, clamped
// taps bypass the box, non-finite tiles and MEMC_B200_FLOAT_ACCUM take fp32 shared atomics).
#include "filter_interpolation.cuh"
#include "tma_utils.cuh"

namespace memc {

namespace {

constexpr int TW = 32;      // tile width; warp w owns tile row w as 4 groups of 8 pixels
constexpr int GW = 8;       // pixels per group: lane = (p = lane & 7, j = lane >> 3)
constexpr int SW = 72;      // box pitch (words)
constexpr int SLAB_H = 4;   // the box is made of slabs of 4 rows

// C channels, TH tile rows (= warps), NSLAB slabs at most, MINB resident CTAs per SM; CACHE: the lane's filter taps and
// gradoutput values of all 4 groups stay in registers between the bound pass and the scatter (28 registers), else they are
// read from shared memory a second time (fewer registers: one more CTA per SM)
template <int C, int TH_, int NSLAB_, int MINB_, bool CACHE_ = true, bool ZTMA_ = false>
struct Lay {
    static constexpr int TH = TH_, NSLAB = NSLAB_, MINB = MINB_, NT = 32 * TH_;
    static constexpr bool CACHE = CACHE_;
    static constexpr bool ZTMA = ZTMA_;  // accumulation slabs zeroed by out-of-bounds TMA loads instead of stores
    static constexpr int STRIP = TH * 16 * GW;  // floats per filter strip [y][i][j][x]
    static constexpr int CH = SLAB_H * SW;      // channel stride inside a slab (words) = 288 = 0 mod 32
    static constexpr int SLAB = C * CH;         // words per slab [c][4][72]
    static constexpr int OFF_GOUT = 4 * STRIP * 4;
    static constexpr int OFF_FLOW = OFF_GOUT + C * TH * TW * 4;
    static constexpr int OFF_BAR = OFF_FLOW + 2 * TH * TW * 4;
    static constexpr int OFF_IMG = OFF_BAR + 128;
    static constexpr int OFF_ACC = OFF_IMG + NSLAB * SLAB * 4;
    static constexpr int TOTAL = OFF_ACC + NSLAB * SLAB * 4;
    static_assert((SLAB * 4) % 128 == 0, "TMA destinations are 128-byte aligned");
};

__device__ __forceinline__ int box_off(int r, int col, int slab_words) {  // row r, column col of channel 0
    return (r >> 2) * slab_words + (r & 3) * SW + col;
}

// Geometry of a pixel as the (pixel, tap row) lanes need it, computed ONCE per pixel by the lane that owns the
// pixel in the row-segment view (lane = x) and handed to the pixel's 4 tap-row lanes with shuffles:
//   code >= 0   fast: window inside the staged box and the image, code = (ly << 8) | lx (box coordinates of tap (0,0))
//   code == -1  valid, but the window touches the image border or leaves the box: per-tap path
//   code == -2  invalid flow / outside the image: contributes nothing (my_lib_kernel.cu:1256)
struct PxGeo {
    int code, ix, iy;
    float alpha, beta;
};

template <class Y, int C, bool OVERWRITE, bool INT_ACC>
__device__ __forceinline__ void compute_rows(const FiArgs& p, float* s_filt, const float* s_gout, float* s_flow, const float* s_img,
                                             float* s_acc, const PxGeo& me, const float (&wt_c)[4][4], const float (&go_c)[4][C], float scale,
                                             int x0, int y0, int b, int bx, int by, int box_rows, int lane, int warp) {
    constexpr int TH = Y::TH, STRIP = Y::STRIP;
    const int W = p.W, H = p.H;
    const int pl = lane & 7, j = lane >> 3;
    const int y = y0 + warp;
    int* s_acci = reinterpret_cast<int*>(s_acc);
    const float* in1b = p.in1p + b * p.in1.b;
    float* g1b = p.gi1p + b * p.gi1.b;
    const bool top = j < 2;
#pragma unroll  // wt[g][] / go[g][] live in registers: g must be a compile-time index
    for (int g = 0; g < 4; ++g) {
        const int src = GW * g + pl, xl = src, x = x0 + xl;
        float wt[4][4], go[4][C];  // only [g] is used: the cached copy, or this group's words read again
        if (Y::CACHE) {
#pragma unroll
            for (int i = 0; i < 4; ++i) wt[g][i] = wt_c[g][i];
#pragma unroll
            for (int c = 0; c < C; ++c) go[g][c] = go_c[g][c];
        } else {
            const float* f = s_filt + g * STRIP + (warp * 16 + j) * GW + pl;
#pragma unroll
            for (int i = 0; i < 4; ++i) wt[g][i] = f[i * 4 * GW];
#pragma unroll
            for (int c = 0; c < C; ++c) go[g][c] = s_gout[(c * TH + warp) * TW + GW * g + pl];
        }
        const int code = __shfl_sync(0xffffffffu, me.code, src);
        const float a = __shfl_sync(0xffffffffu, me.alpha, src), bt = __shfl_sync(0xffffffffu, me.beta, src);
        int Lc = 0, T = 0;
        if (__builtin_expect(__any_sync(0xffffffffu, code == -1), 0)) {  // rare: someone needs full coordinates
            Lc = __shfl_sync(0xffffffffu, me.ix, src) - 1;
            T = __shfl_sync(0xffffffffu, me.iy, src) - 1;
        }
        float a3[4] = {0.f, 0.f, 0.f, 0.f};
        float ql[C], qr[C];
#pragma unroll
        for (int c = 0; c < C; ++c) ql[c] = qr[c] = 0.f;
        if (code != -2) {
            const float wy = top ? (1.0f - bt) : bt;
            float gl[C], gr[C], gls[C], grs[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                gl[c] = go[g][c] * (1.0f - a) * wy;  // gradoutput x bilinear weight of this row's left / right quadrant
                gr[c] = go[g][c] * a * wy;
                gls[c] = gl[c] * scale;              // fixed-point copies (scale is a power of two: exact)
                grs[c] = gr[c] * scale;
            }
            if (__builtin_expect(code >= 0, 1)) {
                const int off = box_off((code >> 8) + j, code & 255, Y::SLAB);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float w = wt[g][i];
                    float acc3 = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int o = off + c * Y::CH + i;
                        const float v = s_img[o];
                        if (INT_ACC) atomicAdd(&s_acci[o], __float2int_rn((i < 2 ? gls[c] : grs[c]) * w));
                        else atomicAdd(&s_acc[o], (i < 2 ? gl[c] : gr[c]) * w);
                        acc3 = fmaf(i < 2 ? gl[c] : gr[c], v, acc3);
                        if (i < 2) ql[c] = fmaf(v, w, ql[c]);
                        else qr[c] = fmaf(v, w, qr[c]);
                    }
                    a3[i] = acc3;
                }
            } else {
                // window touches the image border or leaves the staged box: per-tap clamping; a clamped tap
                // (several can pile up on one border cell) and a tap outside the box go straight to global
                const int cy = clampi(T + j, 0, H - 1);
                const int uy = cy - by;
                const bool row_in = (unsigned)uy < (unsigned)box_rows;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int cx = clampi(Lc + i, 0, W - 1);
                    const int ux = cx - bx;
                    const bool in_box = row_in && (unsigned)ux < (unsigned)SW;
                    const bool to_box = in_box && cx == Lc + i && cy == T + j;
                    const int o0 = box_off(in_box ? uy : 0, in_box ? ux : 0, Y::SLAB);
                    const float w = wt[g][i];
                    float acc3 = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float v = in_box ? s_img[o0 + c * Y::CH] : __ldg(in1b + c * p.in1.c + (int64_t)cy * p.in1.h + cx);
                        const float gs = i < 2 ? gl[c] : gr[c];
                        if (to_box) {
                            if (INT_ACC) atomicAdd(&s_acci[o0 + c * Y::CH], __float2int_rn((i < 2 ? gls[c] : grs[c]) * w));
                            else atomicAdd(&s_acc[o0 + c * Y::CH], gs * w);
                        } else {
                            red_add(g1b + c * p.gi1.c + (int64_t)cy * p.gi1.h + cx, gs * w);
                        }
                        acc3 = fmaf(gs, v, acc3);
                        if (i < 2) ql[c] = fmaf(v, w, ql[c]);
                        else qr[c] = fmaf(v, w, qr[c]);
                    }
                    a3[i] = acc3;
                }
            }
        }
        // ---- flow gradient (the reference's gamma = 1 - beta / 1 - alpha, my_lib_kernel.cu:1358-1495):
        //   d/dx = sum_c go (gam_y (TR - TL) + (1 - gam_y)(BR - BL)),  d/dy = sum_c go (gam_x (BL - TL) + (1 - gam_x)(BR - TR))
        // The 4 tap rows of a pixel sit in lanes p, p + 8, p + 16, p + 24.
        const float gam_y = 1.0f - bt, gam_x = 1.0f - a;
        float dx = 0.f, dy = 0.f;
        if (INT_ACC) {
            // finite tile: every row adds its share (two shuffles per component)
            const float wyg = top ? gam_y : (1.0f - gam_y);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                dx = fmaf(go[g][c] * wyg, qr[c] - ql[c], dx);
                dy = fmaf(go[g][c], gam_x * ql[c] + (1.0f - gam_x) * qr[c], dy);
            }
            dy = top ? -dy : dy;
            dx += __shfl_xor_sync(0xffffffffu, dx, 8);
            dy += __shfl_xor_sync(0xffffffffu, dy, 8);
            dx += __shfl_xor_sync(0xffffffffu, dx, 16);
            dy += __shfl_xor_sync(0xffffffffu, dy, 16);
        } else {
            // non-finite gradients / MEMC_B200_FLOAT_ACCUM: the quadrant sums first, then the reference's expression
            // as written, so that Inf / NaN land exactly where the reference puts them (+Inf - Inf must not appear)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float hl = ql[c] + __shfl_xor_sync(0xffffffffu, ql[c], 8);   // TL (top lanes) / BL (bottom lanes)
                const float hr = qr[c] + __shfl_xor_sync(0xffffffffu, qr[c], 8);   // TR / BR
                const float ol = __shfl_xor_sync(0xffffffffu, hl, 16), orr = __shfl_xor_sync(0xffffffffu, hr, 16);
                const float TL = top ? hl : ol, TR = top ? hr : orr, BL = top ? ol : hl, BR = top ? orr : hr;
                dx = fmaf(go[g][c], gam_y * (TR - TL) + (1.0f - gam_y) * (BR - BL), dx);
                dy = fmaf(go[g][c], gam_x * (BL - TL) + (1.0f - gam_x) * (BR - TR), dy);
            }
        }
        if (code == -2) dx = dy = 0.f;  // (alpha / beta of an invalid pixel may be NaN)
        // gradinput3 of taps (j, 0..3): staged over the filter words this lane read into wt[g][] earlier
        // (all zero for an invalid pixel: stored as such with OVERWRITE, added as such under the += contract)
        {
            float* f = s_filt + g * STRIP + (warp * 16 + j) * GW + pl;
#pragma unroll
            for (int i = 0; i < 4; ++i) f[i * 4 * GW] = a3[i];
        }
        if (OVERWRITE) {
            if (j == 0) {  // nobody reads the flow tile any more (geometry lives in registers)
                s_flow[warp * TW + xl] = dx;
                s_flow[(TH + warp) * TW + xl] = dy;
            }
        } else if (j == 0 && code != -2) {  // assigned for valid pixels only (my_lib_kernel.cu:1424,1495)
            float* g2 = p.gi2p + b * p.gi2.b + (int64_t)y * p.gi2.h + x;
            g2[0] = dx;
            g2[p.gi2.c] = dy;
        }
    }
}

template <class Y, int C, bool OVERWRITE>
__global__ void __launch_bounds__(Y::NT, Y::MINB)
fi_bwd_rows_kernel(const __grid_constant__ CUtensorMap m_flow, const __grid_constant__ CUtensorMap m_gout,
                   const __grid_constant__ CUtensorMap m_filt, const __grid_constant__ CUtensorMap m_img,
                   const __grid_constant__ CUtensorMap m_gi1, const __grid_constant__ CUtensorMap m_gi2,
                   const __grid_constant__ CUtensorMap m_gi3, const __grid_constant__ FiArgs p) {
    constexpr int TH = Y::TH, NT = Y::NT, NSLAB = Y::NSLAB, STRIP = Y::STRIP;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((128u - (tma::smem_u32(smem_raw) & 127u)) & 127u);
    float* s_filt = reinterpret_cast<float*>(sm);                       // 4 x [TH][4 i][4 j][8]
    const float* s_gout = reinterpret_cast<const float*>(sm + Y::OFF_GOUT);  // [C][TH][TW]
    float* s_flow = reinterpret_cast<float*>(sm + Y::OFF_FLOW);         // [2][TH][TW]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Y::OFF_BAR);      // 0 flow, 1 gout, 2 filter, 3 image
    int* s_bb = reinterpret_cast<int*>(bars + 4);
    unsigned* s_max = reinterpret_cast<unsigned*>(s_bb + 4);            // bits of max |gradoutput * filter tap|
    const float* s_img = reinterpret_cast<const float*>(sm + Y::OFF_IMG);
    float* s_acc = reinterpret_cast<float*>(sm + Y::OFF_ACC);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
    const int W = p.W, H = p.H;
    const int pl = lane & 7, j = lane >> 3;

    if (tid == 0) {
        for (int k = 0; k < 4; ++k) tma::mbar_init(&bars[k], 1);
        s_bb[0] = INT_MAX; s_bb[1] = INT_MIN; s_bb[2] = INT_MAX; s_bb[3] = INT_MIN;
        s_max[0] = 0u;
        tma::fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        tma::mbar_expect_tx(&bars[0], 2 * TH * TW * 4);
        tma::load_4d(sm + Y::OFF_FLOW, &m_flow, x0, y0, 0, b, &bars[0]);
        tma::mbar_expect_tx(&bars[1], C * TH * TW * 4);
        tma::load_4d(sm + Y::OFF_GOUT, &m_gout, x0, y0, 0, b, &bars[1]);
        tma::mbar_expect_tx(&bars[2], 4 * STRIP * 4);
#pragma unroll
        for (int g = 0; g < 4; ++g) tma::load_5d(s_filt + g * STRIP, &m_filt, x0 + GW * g, 0, 0, y0, b, &bars[2]);
    }

    // ---- geometry, once per pixel (row-segment view: lane = x), and the bounding box of the source windows
    tma::mbar_wait(&bars[0], 0, 21);
    PxGeo me;
    bool me_valid;
    {
        const FiGeom geo = fi_geometry(x0 + lane, y0 + warp, W, H, s_flow[warp * TW + lane], s_flow[(TH + warp) * TW + lane]);
        me_valid = geo.valid && x0 + lane < W && y0 + warp < H;
        me.ix = geo.ix; me.iy = geo.iy; me.alpha = geo.alpha; me.beta = geo.beta;
        const int mnx = __reduce_min_sync(0xffffffffu, me_valid ? geo.ix : INT_MAX);
        const int mxx = __reduce_max_sync(0xffffffffu, me_valid ? geo.ix : INT_MIN);
        const int mny = __reduce_min_sync(0xffffffffu, me_valid ? geo.iy : INT_MAX);
        const int mxy = __reduce_max_sync(0xffffffffu, me_valid ? geo.iy : INT_MIN);
        if (lane == 0 && mnx <= mxx) {
            atomicMin(&s_bb[0], mnx); atomicMax(&s_bb[1], mxx);
            atomicMin(&s_bb[2], mny); atomicMax(&s_bb[3], mxy);
        }
    }
    __syncthreads();
    const bool any_valid = s_bb[0] <= s_bb[1];
    int bx = 0, by = 0, nslab = 0;
    if (any_valid) {
        // windows span [min ix - 1, max ix + 2] x [min iy - 1, max iy + 2]; a span larger than the box centres it
        // (what it misses takes the per-tap path); x origin rounded down to 4 pixels (TMA: 16-byte coordinates).
        // The box is kept INSIDE the image: a TMA reduce-add at negative coordinates is an illegal instruction on
        // sm_100a (tools/tma_probe.cu, tests 5 and 8), unlike loads, which zero-fill.
        const int need_w = s_bb[1] - s_bb[0] + 4 + 3, need_h = s_bb[3] - s_bb[2] + 4;
        nslab = min(min(NSLAB, H / SLAB_H), (need_h + SLAB_H - 1) / SLAB_H);  // launch precondition: H >= SLAB_H
        bx = s_bb[0] - 1;
        by = s_bb[2] - 1;
        if (need_w > SW) bx += (need_w - SW) / 2;
        if (need_h > nslab * SLAB_H) by += (need_h - nslab * SLAB_H) / 2;
        bx = max(0, min(bx, W - SW)) & ~3;  // W >= SW and W % 4 == 0 are launch preconditions
        by = max(0, min(by, H - nslab * SLAB_H));
    }
    const int box_rows = nslab * SLAB_H;
    if (tid == 0 && any_valid) {
        tma::mbar_expect_tx(&bars[3], (Y::ZTMA ? 2 : 1) * nslab * Y::SLAB * 4);
        for (int s = 0; s < nslab; ++s)
            tma::load_4d(sm + Y::OFF_IMG + s * Y::SLAB * 4, &m_img, bx, by + SLAB_H * s, 0, b, &bars[3]);
        if (Y::ZTMA)  // a box entirely right of the image: the TMA fills it with zeros, no memory traffic, no LSU work
            for (int s = 0; s < nslab; ++s) tma::load_4d(sm + Y::OFF_ACC + s * Y::SLAB * 4, &m_img, W + 64, by, 0, b, &bars[3]);
    }
    {
        const int Lc = me.ix - 1, T = me.iy - 1, lx = Lc - bx, ly = T - by;
        const bool fast = (unsigned)lx <= (unsigned)(SW - 4) && ly >= 0 && ly + 3 < box_rows && Lc >= 0 && Lc + 3 <= W - 1 &&
                          T >= 0 && T + 3 <= H - 1;
        me.code = !me_valid ? -2 : fast ? ((ly << 8) | lx) : -1;
    }
    // zero the slabs in use while the image flies (int 0 and float 0 share the bit pattern)
    if (!Y::ZTMA) {
        float4* a4 = reinterpret_cast<float4*>(s_acc);
        for (int i = tid; i < nslab * Y::SLAB / 4; i += NT) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // ---- this lane's filter taps (row j of 4 pixels) and gradoutput values: read once, kept in registers
    tma::mbar_wait(&bars[1], 0, 22);
    tma::mbar_wait(&bars[2], 0, 23);
    float wt[4][4], go[4][C];
    unsigned mbits = 0u;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const float* f = s_filt + g * STRIP + (warp * 16 + j) * GW + pl;
        unsigned mw = 0u, mg = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            wt[g][i] = f[i * 4 * GW];
            mw = max(mw, __float_as_uint(wt[g][i]) & 0x7fffffffu);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            go[g][c] = s_gout[(c * TH + warp) * TW + GW * g + pl];
            mg = max(mg, __float_as_uint(go[g][c]) & 0x7fffffffu);
        }
        // |x| of a float orders like its bit pattern and NaN patterns sort above +Inf: integer maxima, NaN propagates;
        // Inf * 0 = NaN is fine, any non-finite tile takes the float path.  The maximum over the 4 tap-row lanes of a
        // pixel is the pixel's max_c |gradoutput| * max_t |filter tap|: >= every contribution of that pixel.
        mbits = max(mbits, __float_as_uint(__uint_as_float(mg) * __uint_as_float(mw)) & 0x7fffffffu);
    }
    mbits = __reduce_max_sync(0xffffffffu, mbits);
    if (lane == 0 && mbits) atomicMax(&s_max[0], mbits);
    __syncthreads();  // maximum complete; accumulation slabs zeroed
    float scale = 0.f, inv_scale = 0.f;
    {
        const unsigned mb = s_max[0];
        const float M = __uint_as_float(mb);
        if (mb < 0x7f800000u && M > 0.f && !(p.flags & MEMC_B200_FLOAT_ACCUM)) {
            int ex;
            frexpf(M, &ex);  // M < 2^ex
            constexpr int LOG2_PX = 31 - __builtin_clz(TW * TH - 1) + 1;  // a cell gets <= 1 unclamped tap per pixel
            const int e = max(-120, min(31 - ex - LOG2_PX, 120));
            scale = ldexpf(1.0f, e);
            inv_scale = ldexpf(1.0f, -e);
        }
    }
    if (any_valid) tma::mbar_wait(&bars[3], 0, 24);

    if (scale > 0.f)
        compute_rows<Y, C, OVERWRITE, true>(p, s_filt, s_gout, s_flow, s_img, s_acc, me, wt, go, scale, x0, y0, b, bx, by, box_rows, lane, warp);
    else
        compute_rows<Y, C, OVERWRITE, false>(p, s_filt, s_gout, s_flow, s_img, s_acc, me, wt, go, 1.0f, x0, y0, b, bx, by, box_rows, lane, warp);

    // ---- flush: gradinput3 strips and the gradinput2 tile by TMA store -- issued first, they run while the threads convert
    // the accumulation slabs -- then the slabs of gradinput1 by TMA reduce-add
    tma::fence_proxy_async();  // this thread's staged gradinput3 / gradinput2 words -> visible to the async proxy
    __syncthreads();
    if (tid == 0) {
        // stores / reductions that hang over the right or bottom image edge are clipped by the TMA (tma_probe tests 9-11)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (OVERWRITE) tma::store_5d(&m_gi3, x0 + GW * g, 0, 0, y0, b, s_filt + g * STRIP);
            else tma::reduce_add_5d(&m_gi3, x0 + GW * g, 0, 0, y0, b, s_filt + g * STRIP);
        }
        if (OVERWRITE) tma::store_4d(&m_gi2, x0, y0, 0, b, s_flow);
        tma::bulk_commit();
    }
    if (scale > 0.f) {  // fixed point -> fp32 in place
        int4* ai = reinterpret_cast<int4*>(s_acc);
        float4* af = reinterpret_cast<float4*>(s_acc);
        for (int i = tid; i < nslab * Y::SLAB / 4; i += NT) {
            const int4 q = ai[i];
            af[i] = make_float4((float)q.x * inv_scale, (float)q.y * inv_scale, (float)q.z * inv_scale, (float)q.w * inv_scale);
        }
    }
    tma::fence_proxy_async();  // converted slabs -> visible to the async proxy
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < nslab; ++s) tma::reduce_add_4d(&m_gi1, bx, by + SLAB_H * s, 0, b, s_acc + s * Y::SLAB);
        tma::bulk_commit();
        tma::bulk_wait_read_all();  // shared memory must stay alive until the TMA has read it
    }
}

template <class Y, int C, bool OW>
int launch_rows(cudaStream_t stream, const FiArgs& a) {
    constexpr int TH = Y::TH;
    CUtensorMap m[7];
    const CUtensorMapL2promotion p128 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B, p256 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 pnone = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    if (!tma::make_map_nchw(&m[0], a.flowp, a.B, 2, a.H, a.W, a.flow.b, a.flow.c, a.flow.h, TW, TH, 2, p128) ||
        !tma::make_map_nchw(&m[1], a.goutp, a.B, a.C, a.H, a.W, a.out.b, a.out.c, a.out.h, TW, TH, a.C, p128) ||
        !tma::make_map_taps(&m[2], a.filtp, a.B, a.H, a.W, a.filt.b, a.filt.c, a.filt.h, GW, TH, p256) ||
        !tma::make_map_nchw(&m[3], a.in1p, a.B, a.C, a.H, a.W, a.in1.b, a.in1.c, a.in1.h, SW, SLAB_H, a.C, p128) ||
        !tma::make_map_nchw(&m[4], a.gi1p, a.B, a.C, a.H, a.W, a.gi1.b, a.gi1.c, a.gi1.h, SW, SLAB_H, a.C, pnone) ||
        !tma::make_map_nchw(&m[5], a.gi2p, a.B, 2, a.H, a.W, a.gi2.b, a.gi2.c, a.gi2.h, TW, TH, 2, pnone) ||
        !tma::make_map_taps(&m[6], a.gi3p, a.B, a.H, a.W, a.gi3.b, a.gi3.c, a.gi3.h, GW, TH, pnone))
        return 0;
    constexpr size_t smem = (size_t)Y::TOTAL + 128;
    if (!ensure_dynamic_smem(fi_bwd_rows_kernel<Y, C, OW>, smem)) return 0;
    dim3 grid((a.W + TW - 1) / TW, (a.H + TH - 1) / TH, a.B);
    fi_bwd_rows_kernel<Y, C, OW><<<grid, Y::NT, smem, stream>>>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], a);
    count_launch();
    return check_launch("FilterInterpolation backward (TMA, tap-row lanes)") == 0 ? 1 : -1;
}

template <int C>
int launch_rows_c(cudaStream_t stream, const FiArgs& a, bool ow, int variant) {
    //                       TH NSLAB MINB
    using Y8 = Lay<C, 8, 7, 3, true, true>;  // 32x8 tile, 256 threads, 70 KB (C = 3): 3 CTAs / SM; slabs zeroed by the TMA
    // (0.5172 against 0.5192 ms with the slabs zeroed by stores: the stores ran in the shadow of the image load anyway)
    using Y6 = Lay<C, 6, 5, 4>;  // 32x6 tile, 192 threads, 50 KB: 4 CTAs / SM
    using Y4 = Lay<C, 4, 4, 5>;  // 32x4 tile, 128 threads, 38 KB: 5 CTAs / SM
    using Y12 = Lay<C, 12, 7, 2>;  // 32x12 tile, 384 threads, 80 KB: 2 CTAs / SM
    // (Lay<C, 8, 5, 4, false> -- taps re-read instead of cached, 56 KB, 64 registers, 4 CTAs / SM -- measured 0.562 ms
    // against 0.519 ms: the extra shared-memory reads cost more than the fourth CTA brings; not instantiated)
    if (variant == 4) return ow ? launch_rows<Y12, C, true>(stream, a) : launch_rows<Y12, C, false>(stream, a);
    if (variant == 2) return ow ? launch_rows<Y6, C, true>(stream, a) : launch_rows<Y6, C, false>(stream, a);
    if (variant == 3) return ow ? launch_rows<Y4, C, true>(stream, a) : launch_rows<Y4, C, false>(stream, a);
    return ow ? launch_rows<Y8, C, true>(stream, a) : launch_rows<Y8, C, false>(stream, a);
}

}  // namespace

// 1 = handled, 0 = layout preconditions not met (caller falls back), -1 = launch error
int fi_backward_rows(cudaStream_t stream, const FiArgs& a, bool ow) {
    if (a.fs != 4 || a.C < 1 || a.C > 4 || a.W % 4 || a.B > 65535 || a.W < SW || a.H < SLAB_H) return 0;
    const int variant = (a.flags >> 16) & 0xff;
    switch (a.C) {
        case 1: return launch_rows_c<1>(stream, a, ow, 0);
        case 2: return launch_rows_c<2>(stream, a, ow, 0);
        case 3: return launch_rows_c<3>(stream, a, ow, variant);
        case 4: return launch_rows_c<4>(stream, a, ow, 0);
    }
    return 0;
}

}  // namespace memc
