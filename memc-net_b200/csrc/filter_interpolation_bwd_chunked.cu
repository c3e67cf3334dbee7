// filter_interpolation_bwd_chunked.cu -- FilterInterpolation backward for sm_100a, fs = 4, C > 4 (C % 4 == 0):
// the context-feature warps of MEMC_Net_star (64 channels, networks/MEMC_Net_star.py:280-285) when they are trained.
// Reference semantics my_lib_kernel.cu:1220-1518.  Before this kernel a C > 4 backward ran the generic kernel
// (global float atomics, 5.9 ms per 1080p frame at C = 64).
//
// Same lane map as the C <= 4 production backward (filter_interpolation_bwd_rows.cu): lane = (pixel p, tap ROW j),
// 8x1 pixel groups, box pitch 72, filter tile through the tap-split rank-5 map.  What changes is the channel axis:
//
//   * the tile's geometry, bounding box and the lane's 4 x 4 filter taps are computed / read ONCE and stay in registers;
//   * channels pass through in CHUNKS of 4: a chunk's image box and gradoutput tile arrive by TMA into one of two ring
//     slots (the load of chunk k + 2 is issued as soon as chunk k has left its slot), its gradinput1 contributions are
//     accumulated in ONE int32 box that is converted into the slot the chunk's image came in (dead by then), zeroed
//     in the same pass and leaves through TMA reduce-add -- 3 boxes of shared memory instead of 4;
//   * the filter gradients (4 per lane and group: the taps of row j) and the lane's share of the flow gradient are
//     accumulated in registers ACROSS the chunks and staged / stored once per tile, as in the C <= 4 kernel;
//   * the fixed-point scale is per chunk: the bound max|gradoutput| * max|tap| of chunk k + 1 is taken while chunk k
//     is converted (its gradoutput tile has long arrived), so it costs no extra block barrier: 2 per chunk.
//
// A chunk with non-finite gradients (or MEMC_B200_FLOAT_ACCUM) takes fp32 shared atomics and the reference's flow-
// gradient expression as written, like the C <= 4 kernel.
#include "filter_interpolation.cuh"
#include "tma_utils.cuh"

namespace memc {

namespace {

constexpr int TW = 32;      // tile width; warp w owns tile row w as 4 groups of 8 pixels
constexpr int TH = 8;       // tile rows = warps
constexpr int NT = 32 * TH;
constexpr int GW = 8;       // pixels per group: lane = (p = lane & 7, j = lane >> 3)
constexpr int SW = 72;      // box pitch (words)
constexpr int SLAB_H = 4;   // the box is made of slabs of 4 rows
constexpr int CBK = 4;      // channels per chunk
constexpr int NSLAB = 6;    // 24 box rows at most (what a tile misses takes the per-tap path)

struct Lay {
    static constexpr int STRIP = TH * 16 * GW;  // floats per filter strip [y][i][j][x]
    static constexpr int CH = SLAB_H * SW;      // channel stride inside a slab (words)
    static constexpr int SLAB = CBK * CH;       // words per slab [c][4][72]
    static constexpr int BOX = NSLAB * SLAB;    // words per box
    static constexpr int GOUT = CBK * TH * TW;  // words per gradoutput chunk tile
    static constexpr int OFF_GOUT = 4 * STRIP * 4;
    static constexpr int OFF_FLOW = OFF_GOUT + 2 * GOUT * 4;
    static constexpr int OFF_BAR = OFF_FLOW + 2 * TH * TW * 4;
    static constexpr int OFF_IMG = OFF_BAR + 128;
    static constexpr int OFF_ACC = OFF_IMG + 2 * BOX * 4;
    static constexpr int TOTAL = OFF_ACC + BOX * 4;
    static_assert((SLAB * 4) % 128 == 0 && (GOUT * 4) % 128 == 0, "TMA destinations are 128-byte aligned");
};

__device__ __forceinline__ int box_off(int r, int col) {  // row r, column col of channel 0
    return (r >> 2) * Lay::SLAB + (r & 3) * SW + col;
}

// (see filter_interpolation_bwd_rows.cu)  code >= 0: (ly << 8) | lx, window inside the box and the image; -1: per-tap
// path; -2: invalid pixel, contributes nothing (my_lib_kernel.cu:1256)
struct PxGeo {
    int code, ix, iy;
    float alpha, beta;
};

// one chunk of CBK channels for the 4 pixel groups of this warp's tile row: gradinput1 contributions into the
// accumulation box (or straight to global for clamped / out-of-box taps), filter gradients into a3, flow-gradient
// shares into dxs / dys (per-lane partial sums; the top lanes' dy share is negated at the end of the tile)
template <bool INT_ACC>
__device__ __forceinline__ void chunk_rows(const FiArgs& p, const float* s_img, float* s_acc, const PxGeo& me,
                                           const float (&wt)[4][4], const float (&go)[4][CBK], float scale,
                                           float (&a3)[4][4], float (&dxs)[4], float (&dys)[4], int c0, int b, int bx,
                                           int by, int box_rows, int lane) {
    const int W = p.W, H = p.H;
    const int pl = lane & 7, j = lane >> 3;
    int* s_acci = reinterpret_cast<int*>(s_acc);
    const float* in1b = p.in1p + b * p.in1.b + (int64_t)c0 * p.in1.c;
    float* g1b = p.gi1p + b * p.gi1.b + (int64_t)c0 * p.gi1.c;
    const bool top = j < 2;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int src = GW * g + pl;
        const int code = __shfl_sync(0xffffffffu, me.code, src);
        const float a = __shfl_sync(0xffffffffu, me.alpha, src), bt = __shfl_sync(0xffffffffu, me.beta, src);
        int Lc = 0, T = 0;
        if (__builtin_expect(__any_sync(0xffffffffu, code == -1), 0)) {
            Lc = __shfl_sync(0xffffffffu, me.ix, src) - 1;
            T = __shfl_sync(0xffffffffu, me.iy, src) - 1;
        }
        float ql[CBK], qr[CBK];
#pragma unroll
        for (int c = 0; c < CBK; ++c) ql[c] = qr[c] = 0.f;
        if (code != -2) {
            const float wy = top ? (1.0f - bt) : bt;
            float gl[CBK], gr[CBK];
#pragma unroll
            for (int c = 0; c < CBK; ++c) {
                gl[c] = go[g][c] * (1.0f - a) * wy;  // gradoutput x bilinear weight of this row's left / right quadrant
                gr[c] = go[g][c] * a * wy;
            }
            if (__builtin_expect(code >= 0, 1)) {
                const int off = box_off((code >> 8) + j, code & 255);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float w = wt[g][i];
                    const float ws = w * scale;  // scale is a power of two: exact
                    float acc3 = a3[g][i];
#pragma unroll
                    for (int c = 0; c < CBK; ++c) {
                        const int o = off + c * Lay::CH + i;
                        const float v = s_img[o];
                        const float gq = i < 2 ? gl[c] : gr[c];
                        if (INT_ACC) atomicAdd(&s_acci[o], __float2int_rn(gq * ws));
                        else atomicAdd(&s_acc[o], gq * w);
                        acc3 = fmaf(gq, v, acc3);
                        if (i < 2) ql[c] = fmaf(v, w, ql[c]);
                        else qr[c] = fmaf(v, w, qr[c]);
                    }
                    a3[g][i] = acc3;
                }
            } else {
                // window touches the image border or leaves the staged box: per-tap clamping; a clamped tap and a
                // tap outside the box go straight to global
                const int cy = clampi(T + j, 0, H - 1);
                const int uy = cy - by;
                const bool row_in = (unsigned)uy < (unsigned)box_rows;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int cx = clampi(Lc + i, 0, W - 1);
                    const int ux = cx - bx;
                    const bool in_box = row_in && (unsigned)ux < (unsigned)SW;
                    const bool to_box = in_box && cx == Lc + i && cy == T + j;
                    const int o0 = box_off(in_box ? uy : 0, in_box ? ux : 0);
                    const float w = wt[g][i];
                    float acc3 = a3[g][i];
#pragma unroll
                    for (int c = 0; c < CBK; ++c) {
                        const float v = in_box ? s_img[o0 + c * Lay::CH] : __ldg(in1b + c * p.in1.c + (int64_t)cy * p.in1.h + cx);
                        const float gq = i < 2 ? gl[c] : gr[c];
                        if (to_box) {
                            if (INT_ACC) atomicAdd(&s_acci[o0 + c * Lay::CH], __float2int_rn(gq * (w * scale)));
                            else atomicAdd(&s_acc[o0 + c * Lay::CH], gq * w);
                        } else {
                            red_add(g1b + c * p.gi1.c + (int64_t)cy * p.gi1.h + cx, gq * w);
                        }
                        acc3 = fmaf(gq, v, acc3);
                        if (i < 2) ql[c] = fmaf(v, w, ql[c]);
                        else qr[c] = fmaf(v, w, qr[c]);
                    }
                    a3[g][i] = acc3;
                }
            }
            // ---- flow gradient (my_lib_kernel.cu:1358-1495), see filter_interpolation_bwd_rows.cu
            const float gam_y = 1.0f - bt, gam_x = 1.0f - a;
            if (INT_ACC) {
                const float wyg = top ? gam_y : (1.0f - gam_y);
                float dx = 0.f, dy = 0.f;
#pragma unroll
                for (int c = 0; c < CBK; ++c) {
                    dx = fmaf(go[g][c] * wyg, qr[c] - ql[c], dx);
                    dy = fmaf(go[g][c], gam_x * ql[c] + (1.0f - gam_x) * qr[c], dy);
                }
                dxs[g] += dx;
                dys[g] += dy;
            }
        }
        if (!INT_ACC) {
            // non-finite chunk / MEMC_B200_FLOAT_ACCUM: quadrant sums first, then the reference's expression as written
            // (+Inf - Inf must not appear where the reference has none); the full value goes into the j == 0 lane's
            // share (negated for dy: the top lanes' shares are negated at the end of the tile).  All 32 lanes shuffle.
            const float gam_y = 1.0f - bt, gam_x = 1.0f - a;
            float dx = 0.f, dy = 0.f;
#pragma unroll
            for (int c = 0; c < CBK; ++c) {
                const float hl = ql[c] + __shfl_xor_sync(0xffffffffu, ql[c], 8);
                const float hr = qr[c] + __shfl_xor_sync(0xffffffffu, qr[c], 8);
                const float ol = __shfl_xor_sync(0xffffffffu, hl, 16), orr = __shfl_xor_sync(0xffffffffu, hr, 16);
                const float TL = top ? hl : ol, TR = top ? hr : orr, BL = top ? ol : hl, BR = top ? orr : hr;
                dx = fmaf(go[g][c], gam_y * (TR - TL) + (1.0f - gam_y) * (BR - BL), dx);
                dy = fmaf(go[g][c], gam_x * (BL - TL) + (1.0f - gam_x) * (BR - TR), dy);
            }
            if (j == 0 && code != -2) {
                dxs[g] += dx;
                dys[g] -= dy;
            }
        }
    }
}

template <bool OVERWRITE>
__global__ void __launch_bounds__(NT, 2)
fi_bwd_chunked_kernel(const __grid_constant__ CUtensorMap m_flow, const __grid_constant__ CUtensorMap m_gout,
                      const __grid_constant__ CUtensorMap m_filt, const __grid_constant__ CUtensorMap m_img,
                      const __grid_constant__ CUtensorMap m_gi1, const __grid_constant__ CUtensorMap m_gi2,
                      const __grid_constant__ CUtensorMap m_gi3, const __grid_constant__ FiArgs p) {
    using Y = Lay;
    constexpr int STRIP = Y::STRIP;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((128u - (tma::smem_u32(smem_raw) & 127u)) & 127u);
    float* s_filt = reinterpret_cast<float*>(sm);                          // 4 x [TH][4 i][4 j][8]
    const float* s_gout = reinterpret_cast<const float*>(sm + Y::OFF_GOUT);  // 2 x [CBK][TH][TW]
    float* s_flow = reinterpret_cast<float*>(sm + Y::OFF_FLOW);            // [2][TH][TW]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Y::OFF_BAR);         // 0 flow, 1 filter, 2/3 chunk slot 0/1
    int* s_bb = reinterpret_cast<int*>(bars + 4);
    unsigned* s_max = reinterpret_cast<unsigned*>(s_bb + 4);               // [2]: bits of the chunk's bound
    float* s_img = reinterpret_cast<float*>(sm + Y::OFF_IMG);              // 2 boxes
    float* s_acc = reinterpret_cast<float*>(sm + Y::OFF_ACC);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
    const int W = p.W, H = p.H;
    const int pl = lane & 7, j = lane >> 3;
    const int nchunk = p.C / CBK;

    if (tid == 0) {
        for (int k = 0; k < 4; ++k) tma::mbar_init(&bars[k], 1);
        s_bb[0] = INT_MAX; s_bb[1] = INT_MIN; s_bb[2] = INT_MAX; s_bb[3] = INT_MIN;
        s_max[0] = s_max[1] = 0u;
        tma::fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        tma::mbar_expect_tx(&bars[0], 2 * TH * TW * 4);
        tma::load_4d(sm + Y::OFF_FLOW, &m_flow, x0, y0, 0, b, &bars[0]);
        tma::mbar_expect_tx(&bars[1], 4 * STRIP * 4);
#pragma unroll
        for (int g = 0; g < 4; ++g) tma::load_5d(s_filt + g * STRIP, &m_filt, x0 + GW * g, 0, 0, y0, b, &bars[1]);
    }

    // ---- geometry, once per pixel (row-segment view: lane = x), and the bounding box of the source windows
    tma::mbar_wait(&bars[0], 0, 31);
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
        // as in the C <= 4 kernel: the box is kept inside the image (a TMA reduce-add at negative coordinates is an
        // illegal instruction), x origin on a 16-byte boundary, a span larger than the box centres it
        const int need_w = s_bb[1] - s_bb[0] + 4 + 3, need_h = s_bb[3] - s_bb[2] + 4;
        nslab = min(min(NSLAB, H / SLAB_H), (need_h + SLAB_H - 1) / SLAB_H);
        bx = s_bb[0] - 1;
        by = s_bb[2] - 1;
        if (need_w > SW) bx += (need_w - SW) / 2;
        if (need_h > nslab * SLAB_H) by += (need_h - nslab * SLAB_H) / 2;
        bx = max(0, min(bx, W - SW)) & ~3;
        by = max(0, min(by, H - nslab * SLAB_H));
    }
    const int box_rows = nslab * SLAB_H;
    const int nch = any_valid ? nchunk : 0;  // a tile without a valid pixel has no gradients at all
    // chunk k -> ring slot k & 1: image box slabs + gradoutput tile, one mbarrier phase per use of the slot
    auto issue_chunk = [&](int k) {
        const int s = k & 1;
        tma::mbar_expect_tx(&bars[2 + s], (nslab * Y::SLAB + Y::GOUT) * 4);
        tma::load_4d(sm + Y::OFF_GOUT + s * Y::GOUT * 4, &m_gout, x0, y0, k * CBK, b, &bars[2 + s]);
        for (int q = 0; q < nslab; ++q)
            tma::load_4d(s_img + s * Y::BOX + q * Y::SLAB, &m_img, bx, by + SLAB_H * q, k * CBK, b, &bars[2 + s]);
    };
    if (tid == 0) {
        if (nch > 0) issue_chunk(0);
        if (nch > 1) issue_chunk(1);
    }
    {
        const int Lc = me.ix - 1, T = me.iy - 1, lx = Lc - bx, ly = T - by;
        const bool fast = (unsigned)lx <= (unsigned)(SW - 4) && ly >= 0 && ly + 3 < box_rows && Lc >= 0 && Lc + 3 <= W - 1 &&
                          T >= 0 && T + 3 <= H - 1;
        me.code = !me_valid ? -2 : fast ? ((ly << 8) | lx) : -1;
    }
    {
        float4* a4 = reinterpret_cast<float4*>(s_acc);
        for (int i = tid; i < nslab * Y::SLAB / 4; i += NT) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // ---- this lane's filter taps (row j of 4 pixels): read once, in registers for all chunks
    tma::mbar_wait(&bars[1], 0, 32);
    float wt[4][4];
    float mw[4];  // max |tap| of the lane's row, per group
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const float* f = s_filt + g * STRIP + (warp * 16 + j) * GW + pl;
        unsigned m = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            wt[g][i] = f[i * 4 * GW];
            m = max(m, __float_as_uint(wt[g][i]) & 0x7fffffffu);
        }
        mw[g] = __uint_as_float(m);
    }
    // gradoutput values of a chunk for this lane's 4 pixels, and the chunk's bound (|x| of a float orders like its bit
    // pattern and NaN patterns sort above +Inf: integer maxima, NaN propagates; see the C <= 4 kernel)
    float go[4][CBK];
    auto fetch_go = [&](int k) {
        const int s = k & 1;
        tma::mbar_wait(&bars[2 + s], (k >> 1) & 1, 33);
        const float* gsrc = s_gout + s * Y::GOUT;
        unsigned mbits = 0u;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            unsigned mg = 0u;
#pragma unroll
            for (int c = 0; c < CBK; ++c) {
                go[g][c] = gsrc[(c * TH + warp) * TW + GW * g + pl];
                mg = max(mg, __float_as_uint(go[g][c]) & 0x7fffffffu);
            }
            mbits = max(mbits, __float_as_uint(__uint_as_float(mg) * mw[g]) & 0x7fffffffu);
        }
        mbits = __reduce_max_sync(0xffffffffu, mbits);
        if (lane == 0 && mbits) atomicMax(&s_max[s], mbits);
    };
    if (nch > 0) fetch_go(0);
    __syncthreads();  // bound of chunk 0 complete; accumulation box zeroed

    float a3[4][4], dxs[4], dys[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        dxs[g] = dys[g] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) a3[g][i] = 0.f;
    }

    for (int k = 0; k < nch; ++k) {
        const int s = k & 1;
        float scale = 0.f, inv_scale = 0.f;
        {
            const unsigned mb = s_max[s];
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
        const float* img = s_img + s * Y::BOX;
        if (scale > 0.f) chunk_rows<true>(p, img, s_acc, me, wt, go, scale, a3, dxs, dys, k * CBK, b, bx, by, box_rows, lane);
        else chunk_rows<false>(p, img, s_acc, me, wt, go, 1.0f, a3, dxs, dys, k * CBK, b, bx, by, box_rows, lane);
        __syncthreads();  // every contribution of the chunk is in the box; its image slot is dead
        if (tid == 0) s_max[s] = 0u;  // (read by everyone before the barrier; written again two chunks from now)
        // box -> fp32 into the dead image slot, box zeroed for the next chunk in the same pass
        {
            int4* ai = reinterpret_cast<int4*>(s_acc);
            float4* af = reinterpret_cast<float4*>(s_img + s * Y::BOX);
            if (scale > 0.f) {
                for (int i = tid; i < nslab * Y::SLAB / 4; i += NT) {
                    const int4 q = ai[i];
                    ai[i] = make_int4(0, 0, 0, 0);
                    af[i] = make_float4((float)q.x * inv_scale, (float)q.y * inv_scale, (float)q.z * inv_scale, (float)q.w * inv_scale);
                }
            } else {
                const float4* sf = reinterpret_cast<const float4*>(s_acc);
                for (int i = tid; i < nslab * Y::SLAB / 4; i += NT) {
                    af[i] = sf[i];
                    ai[i] = make_int4(0, 0, 0, 0);
                }
            }
        }
        if (k + 1 < nch) fetch_go(k + 1);  // its tile arrived a chunk ago; the bound is published by the next barrier
        tma::fence_proxy_async();  // converted slabs -> visible to the async proxy
        __syncthreads();
        if (tid == 0) {
            for (int q = 0; q < nslab; ++q)
                tma::reduce_add_4d(&m_gi1, bx, by + SLAB_H * q, k * CBK, b, s_img + s * Y::BOX + q * Y::SLAB);
            tma::bulk_commit();
            if (k + 2 < nch) {
                tma::bulk_wait_read_all();  // the slot has been read: refill it
                issue_chunk(k + 2);
            }
        }
    }

    // ---- gradinput3 (taps (j, 0..3) of 4 pixels per lane) staged over the filter words this lane read; gradinput2
    // reduced over the 4 tap-row lanes of a pixel and staged over the flow tile
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float* f = s_filt + g * STRIP + (warp * 16 + j) * GW + pl;
#pragma unroll
        for (int i = 0; i < 4; ++i) f[i * 4 * GW] = a3[g][i];
        float dx = dxs[g], dy = j < 2 ? -dys[g] : dys[g];
        dx += __shfl_xor_sync(0xffffffffu, dx, 8);
        dy += __shfl_xor_sync(0xffffffffu, dy, 8);
        dx += __shfl_xor_sync(0xffffffffu, dx, 16);
        dy += __shfl_xor_sync(0xffffffffu, dy, 16);
        const int xl = GW * g + pl;
        const int code = __shfl_sync(0xffffffffu, me.code, xl);
        if (OVERWRITE) {
            if (j == 0) {
                s_flow[warp * TW + xl] = code == -2 ? 0.f : dx;
                s_flow[(TH + warp) * TW + xl] = code == -2 ? 0.f : dy;
            }
        } else if (j == 0 && code != -2) {  // assigned for valid pixels only (my_lib_kernel.cu:1424,1495)
            float* g2 = p.gi2p + b * p.gi2.b + (int64_t)(y0 + warp) * p.gi2.h + x0 + xl;
            g2[0] = dx;
            g2[p.gi2.c] = dy;
        }
    }
    tma::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (OVERWRITE) tma::store_5d(&m_gi3, x0 + GW * g, 0, 0, y0, b, s_filt + g * STRIP);
            else tma::reduce_add_5d(&m_gi3, x0 + GW * g, 0, 0, y0, b, s_filt + g * STRIP);
        }
        if (OVERWRITE) tma::store_4d(&m_gi2, x0, y0, 0, b, s_flow);
        tma::bulk_commit();
        tma::bulk_wait_read_all();  // shared memory must stay alive until the TMA has read it
    }
}

template <bool OW>
int launch_chunked(cudaStream_t stream, const FiArgs& a) {
    CUtensorMap m[7];
    const CUtensorMapL2promotion p128 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B, p256 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 pnone = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    if (!tma::make_map_nchw(&m[0], a.flowp, a.B, 2, a.H, a.W, a.flow.b, a.flow.c, a.flow.h, TW, TH, 2, p128) ||
        !tma::make_map_nchw(&m[1], a.goutp, a.B, a.C, a.H, a.W, a.out.b, a.out.c, a.out.h, TW, TH, CBK, p128) ||
        !tma::make_map_taps(&m[2], a.filtp, a.B, a.H, a.W, a.filt.b, a.filt.c, a.filt.h, GW, TH, p256) ||
        !tma::make_map_nchw(&m[3], a.in1p, a.B, a.C, a.H, a.W, a.in1.b, a.in1.c, a.in1.h, SW, SLAB_H, CBK, p128) ||
        !tma::make_map_nchw(&m[4], a.gi1p, a.B, a.C, a.H, a.W, a.gi1.b, a.gi1.c, a.gi1.h, SW, SLAB_H, CBK, pnone) ||
        !tma::make_map_nchw(&m[5], a.gi2p, a.B, 2, a.H, a.W, a.gi2.b, a.gi2.c, a.gi2.h, TW, TH, 2, pnone) ||
        !tma::make_map_taps(&m[6], a.gi3p, a.B, a.H, a.W, a.gi3.b, a.gi3.c, a.gi3.h, GW, TH, pnone))
        return 0;
    constexpr size_t smem = (size_t)Lay::TOTAL + 128;
    if (!ensure_dynamic_smem(fi_bwd_chunked_kernel<OW>, smem)) return 0;
    dim3 grid((a.W + TW - 1) / TW, (a.H + TH - 1) / TH, a.B);
    fi_bwd_chunked_kernel<OW><<<grid, NT, smem, stream>>>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], a);
    count_launch();
    return check_launch("FilterInterpolation backward (TMA, tap-row lanes, channel chunks)") == 0 ? 1 : -1;
}

}  // namespace

// 1 = handled, 0 = layout preconditions not met (caller falls back), -1 = launch error
int fi_backward_chunked(cudaStream_t stream, const FiArgs& a, bool ow) {
    if (a.fs != 4 || a.C <= 4 || a.C % CBK || a.W % 4 || a.B > 65535 || a.W < SW || a.H < SLAB_H) return 0;
    return ow ? launch_chunked<true>(stream, a) : launch_chunked<false>(stream, a);
}

}  // namespace memc
