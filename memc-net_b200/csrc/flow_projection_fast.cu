// flow_projection_fast.cu -- FlowProjection forward for sm_100a: shared-memory privatised splat,
// occupancy-mask hole filling, and (for MEMC_B200_OVERWRITE calls) the whole forward as one
// persistent software-pipelined kernel.
//
// Same arithmetic as flow_projection.cu (reference my_lib_kernel.cu:1630-1836), different data
// movement.  The legacy splat issues 12 global float atomics per source pixel (ncu: ~6.6 L2
// reduction sectors per pixel, L2-atomic bound; 31 ms per 16 frames when flows converge).  Here a
// CTA owns a TW x TH tile of SOURCE pixels:
//
//   LDG   the flow of the tile's pixels, straight into registers (prefetched one tile ahead in
//         the persistent kernel)
//   ...   targets p + flow, bounding box of the target cells, M = max|flow| over the tile
//   ...   the tile's contributions (-fx, -fy, +1) are accumulated in a SHARED-MEMORY box of
//         3 planes (x, y, count) with native int32 shared atomics: count is an integer anyway,
//         -fx / -fy are accumulated in fixed point with a per-tile power-of-two scale
//         2^e, M * (TW*TH) * 2^e < 2^31 -- order independent within the tile.
//         The reference adds the SAME value to the four cells (L..L+1) x (T..T+1), so the box
//         holds the "corner histogram" A[T][L] (3 shared atomics per source, not 12) and the
//         flush applies the 2x2 box filter;
//   ...   the touched part of the box is flushed with 128-bit vector reductions (4 cells each,
//         all-zero vectors skipped).  Targets outside the box fall back to global atomics.
#include "flow_projection.cuh"
#include <limits.h>
#include <mutex>

namespace memc {

namespace {

constexpr int TW = 64, TH = 16, NT = 256, PPT = TW * TH / NT;  // source tile, 4 pixels / thread
constexpr int SW = 96, SH = 32;                                // target box (pitch 96 = 3*32 words)
constexpr int BOX = SW * SH;

// per-tile control words; two sets, used alternately by consecutive splat tiles of a CTA: a tile resets the set of the
// NEXT tile, so that a tile's atomics need no barrier between "control words initialised" and "first atomic"
struct Ctl {
    int bb[4];
    unsigned maxbits;
    int kmax;  // largest number of sources of this tile that share one corner cell
    __device__ __forceinline__ void reset() {
        bb[0] = INT_MAX; bb[1] = INT_MIN; bb[2] = INT_MAX; bb[3] = INT_MIN;
        maxbits = 0u;
        kmax = 0;
    }
};
struct __align__(128) Smem {
    int box[3][SH][SW];  // x, y (fixed point) and count: 36 KB
    Ctl ctl[2];
};

__device__ __forceinline__ bool fp_valid(float x2, float y2, int W, int H) {
    return x2 >= 0.0f && y2 >= 0.0f && x2 <= (float)(W - 1) && y2 <= (float)(H - 1);
}

// pixel k of this thread inside a source tile: a warp owns 32-pixel row segments
__device__ __forceinline__ void tile_pixel(int k, int& xl, int& yl) {
    const int seg = (threadIdx.x >> 5) + k * (NT / 32);
    yl = seg / (TW / 32);
    xl = (threadIdx.x & 31) + 32 * (seg % (TW / 32));
}

// L2 eviction policies (createpolicy): the persistent pipeline marks what it streams (flow in) evict-first and what it
// keeps coming back to (the accumulator ring) evict-last, so that the 75 MB ring is what stays in the 126 MB L2.
// kind: 0 normal, 1 evict-first, 2 evict-last
__device__ __forceinline__ uint64_t l2_policy(int kind) {
    uint64_t pol;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float ldg_stream_pol(const float* p, uint64_t pol) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void red_add4_pol(float* p, float4 v, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ float4 ldcg4_pol(const float* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol)
                 : "memory");
    return v;
}
__device__ __forceinline__ void stcg4_pol(float* p, float4 v, uint64_t pol) {
    asm volatile("st.global.cg.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w), "l"(pol)
                 : "memory");
}

// flow of this thread's pixels of tile (x0, y0) of frame b: coalesced read-only loads
template <bool POL = false>
__device__ __forceinline__ void load_flow(const float* __restrict__ flowp, View fv, int x0, int y0, int b, int W, int H,
                                          float (&fx)[PPT], float (&fy)[PPT], uint64_t pol = 0) {
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        int xl, yl;
        tile_pixel(k, xl, yl);
        const int x = x0 + xl, y = y0 + yl;
        fx[k] = fy[k] = 0.f;
        if (x < W && y < H) {
            const float* f = flowp + (int64_t)b * fv.b + (int64_t)y * fv.h + x;
            fx[k] = POL ? ldg_stream_pol(f, pol) : ldg_stream(f);
            fy[k] = POL ? ldg_stream_pol(f + fv.c, pol) : ldg_stream(f + fv.c);
        }
    }
}

// One source tile.  The reference adds the SAME value (-fx, -fy, 1) to the four cells
// (L..L+1) x (T..T+1) of every valid source pixel (my_lib_kernel.cu:1673-1689).  The sum over
// sources is therefore a 2x2 box filter of the "corner histogram" A[T][L] += value: the tile
// issues 3 shared atomics per source pixel (one cell) instead of 12, and the flush evaluates
//     cell(x, y) = A[y][x] + A[y][x-1] + A[y-1][x] + A[y-1][x-1]
// in exact integer arithmetic (every source is in exactly one A cell, so no partial sum can
// exceed the tile total the fixed-point scale was sized for).  Taps whose cell is clamped at the
// right / bottom image border (R == L or Bm == T) do not follow the box pattern: they go straight
// to global memory, as do sources whose corner cell misses the staged box.
//
// Precondition: a __syncthreads() separates this call from the CTA's previous use of `s`.
// PRE: the control set s.ctl[par] was reset before that barrier (by the previous tile): no barrier is needed before this
// tile's first atomics on it (one barrier less per tile in the persistent pipeline).
template <bool PRE>
__device__ __forceinline__ void splat_tile(Smem& s_, int par, const float (&fx)[PPT], const float (&fy)[PPT], int x0, int y0,
                                           float* ox, float* oy, float* cn, int64_t out_h, int64_t cnt_h, int W, int H,
                                           uint64_t pol = 0) {
    const int tid = threadIdx.x, lane = tid & 31;
    struct View_ { int (&box)[3][SH][SW]; int (&bb)[4]; unsigned& maxbits; int& kmax; };
    View_ s{s_.box, s_.ctl[par].bb, s_.ctl[par].maxbits, s_.ctl[par].kmax};
    if (tid == 0) {
        s_.ctl[par ^ 1].reset();  // for the next tile (nobody touches that set until the barrier at the end of this tile)
        if (!PRE) s_.ctl[par].reset();
    }
    {   // zero the box
        int4* z = reinterpret_cast<int4*>(&s.box[0][0][0]);
        for (int i = tid; i < 3 * BOX / 4; i += NT) z[i] = make_int4(0, 0, 0, 0);
    }
    // ---- targets of my pixels, tile bounding box of the corner cells, tile max |flow|
    int L[PPT], T[PPT];
    bool ok[PPT];
    int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
    float mloc = 0.f;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        int xl, yl;
        tile_pixel(k, xl, yl);
        const int x = x0 + xl, y = y0 + yl;
        const float x2 = (float)x + fx[k], y2 = (float)y + fy[k];
        ok[k] = x < W && y < H && fp_valid(x2, y2, W, H);
        L[k] = ok[k] ? (int)x2 : 0;
        T[k] = ok[k] ? (int)y2 : 0;
        if (ok[k]) {
            mnx = min(mnx, L[k]); mxx = max(mxx, L[k]);
            mny = min(mny, T[k]); mxy = max(mxy, T[k]);
            mloc = fmaxf_nan(mloc, fmaxf_nan(fabsf(fx[k]), fabsf(fy[k])));
        }
    }
    mnx = __reduce_min_sync(0xffffffffu, mnx); mxx = __reduce_max_sync(0xffffffffu, mxx);  // REDUX
    mny = __reduce_min_sync(0xffffffffu, mny); mxy = __reduce_max_sync(0xffffffffu, mxy);
    const unsigned mb = __reduce_max_sync(0xffffffffu, __float_as_uint(mloc));
    if (!PRE) __syncthreads();  // control words initialised
    if (lane == 0 && mnx <= mxx) {
        atomicMin(&s.bb[0], mnx); atomicMax(&s.bb[1], mxx);
        atomicMin(&s.bb[2], mny); atomicMax(&s.bb[3], mxy);
        if (mb) atomicMax(&s.maxbits, mb);
    }
    __syncthreads();  // bounding box and maximum complete; box zeroed
    if (s.bb[0] > s.bb[1]) return;  // no valid source pixel in this tile (uniform across the CTA)

    // corner cells span [min L, max L] x [min T, max T]; box x origin multiple of 4 (vector flush)
    int bx = s.bb[0], by = s.bb[2];
    {
        const int need_w = s.bb[1] - s.bb[0] + 1 + 3, need_h = s.bb[3] - s.bb[2] + 1;
        if (need_w > SW) bx += (need_w - SW) / 2;
        if (need_h > SH) by += (need_h - SH) / 2;
        bx = max(0, min(bx, W - SW)) & ~3;
        by = max(0, min(by, H - SH));
    }
    // ---- count first: the hit counts are integers anyway, and the value the atomic returns tells how
    // many sources of this tile share a corner cell.  K = the largest such multiplicity bounds every
    // corner-cell sum by K * M and every 2x2-filtered output cell by 4 * K * M, so the fixed-point scale
    // can spend the bits a worst-case bound (all TW*TH sources on one cell) would waste: on a smooth
    // field K is 2..4 and a contribution is rounded to ~M * 2^-27 instead of M * 2^-21 -- below the
    // rounding of the reference's own fp32 atomics (tests/test_gpu_at_size.py prints both).
    bool in_box[PPT];
    int kloc = 0;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int ux = L[k] - bx, uy = T[k] - by;
        in_box[k] = ok[k] && (unsigned)ux < (unsigned)SW && (unsigned)uy < (unsigned)SH;
        if (in_box[k]) kloc = max(kloc, atomicAdd(&s.box[2][uy][ux], 1) + 1);
    }
    kloc = __reduce_max_sync(0xffffffffu, kloc);
    if (lane == 0 && kloc) atomicMax(&s.kmax, kloc);
    __syncthreads();
    // fixed-point scale (see header); valid pixels have finite flows, so M is finite
    const float M = __uint_as_float(s.maxbits);
    float scale = 1.0f, inv_scale = 1.0f;
    if (M > 0.f) {
        int ex;
        frexpf(M, &ex);
        constexpr int LOG2_PX = 31 - __builtin_clz(TW * TH - 1) + 1;  // every source is in exactly one cell
        const int log2_4k = 32 - __clz(4 * max(s.kmax, 1) - 1);       // ceil(log2(4 K))
        const int e = max(-120, min(31 - ex - min(LOG2_PX, log2_4k), 120));
        scale = ldexpf(1.0f, e);
        inv_scale = ldexpf(1.0f, -e);
    }

#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        if (!ok[k]) continue;
        const int ux = L[k] - bx, uy = T[k] - by;
        if (in_box[k]) {
            atomicAdd(&s.box[0][uy][ux], __float2int_rn(-fx[k] * scale));
            atomicAdd(&s.box[1][uy][ux], __float2int_rn(-fy[k] * scale));
        }
        const bool last_col = L[k] == W - 1, last_row = T[k] == H - 1;
        if (__builtin_expect(!in_box[k] || last_col || last_row, 0)) {
            const int R = min(L[k] + 1, W - 1), Bm = min(T[k] + 1, H - 1);
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    // in the box: only the clamped repeats are missing from the 2x2 pattern
                    if (in_box[k] && !((i == 1 && last_col) || (j == 1 && last_row))) continue;
                    const int cx = i ? R : L[k], cy = j ? Bm : T[k];
                    red_add(ox + (int64_t)cy * out_h + cx, -fx[k]);
                    red_add(oy + (int64_t)cy * out_h + cx, -fy[k]);
                    red_add(cn + (int64_t)cy * cnt_h + cx, 1.0f);
                }
        }
    }
    __syncthreads();
    // ---- flush the touched corner cells: four cells per 128-bit vector reduction
    // (REDG.E.ADD.F32x4), all-zero vectors skipped; a warp walks rows, its lanes the vectors of a
    // row.  The 2x2 box filter is applied on the way out; output cells beyond the image
    // are dropped (their taps were sent directly above).
    {
        const int ax0 = max(s.bb[0], bx) - bx, ax1 = min(s.bb[1], bx + SW - 1) - bx;
        const int ay0 = max(s.bb[2], by) - by, ay1 = min(s.bb[3], by + SH - 1) - by;
        if (ax0 <= ax1 && ay0 <= ay1) {
            const int cx1 = min(ax1 + 1, W - 1 - bx), cy1 = min(ay1 + 1, H - 1 - by);
            const int v0 = ax0 >> 2, v1 = cx1 >> 2;  // vector columns (ux = 4 v <= SW)
            for (int uy = ay0 + (tid >> 5); uy <= cy1; uy += NT / 32)
                for (int v = v0 + lane; v <= v1; v += 32) {
                    const int ux = v << 2;
#pragma unroll
                    for (int pl = 0; pl < 3; ++pl) {
                        int q[4], a[2][5];
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            const int yy = uy - 1 + rr;
                            const bool row_ok = (unsigned)yy < (unsigned)SH;
                            a[rr][0] = (row_ok && ux > 0) ? s.box[pl][yy][ux - 1] : 0;
                            int4 t = make_int4(0, 0, 0, 0);
                            if (row_ok && ux < SW) t = *reinterpret_cast<const int4*>(&s.box[pl][yy][ux]);
                            a[rr][1] = t.x; a[rr][2] = t.y; a[rr][3] = t.z; a[rr][4] = t.w;
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) q[k] = a[0][k] + a[0][k + 1] + a[1][k] + a[1][k + 1];
                        if ((q[0] | q[1] | q[2] | q[3]) == 0) continue;
                        const float sc = pl == 2 ? 1.0f : inv_scale;
                        const float4 val = make_float4((float)q[0] * sc, (float)q[1] * sc, (float)q[2] * sc, (float)q[3] * sc);
                        float* dst = (pl == 0 ? ox : pl == 1 ? oy : cn) + (int64_t)(by + uy) * (pl == 2 ? cnt_h : out_h) + bx + ux;
                        if (PRE) red_add4_pol(dst, val, pol);  // pipeline: the accumulator ring, L2 evict-last
                        else atomicAdd(reinterpret_cast<float4*>(dst), val);
                    }
                }
        }
    }
}

// ------------------------------------------------------------------------------------
// Reference-contract calls (the caller's pre-zeroed buffers are accumulated into): frame by frame
// memset -> splat -> average + masks, then one fill-hole launch.  A frame's count+output planes
// (25 MB at 1080p) stay L2-resident between the passes.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 4) fp_splat_kernel(const FpArgs p, const int b) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    float fx[PPT], fy[PPT];
    load_flow(p.flowp, p.flow, x0, y0, b, p.W, p.H, fx, fy);
    float* ox = p.outp + (int64_t)b * p.out.b;
    splat_tile<false>(s, 0, fx, fy, x0, y0, ox, ox + p.out.c, p.countp + (int64_t)b * p.count.b, p.out.h, p.count.h, p.W, p.H);
}

// average + occupancy bit masks.  One CTA = one 128 x 32 pixel block:
//   out /= count where count > 0 (my_lib_kernel.cu:1730-1736), and
//   rowmask[b][y][x/32]  bit (x%32) = count[b][y][x] > 0      (32 pixels of a row per word)
//   colmask[b][y/32][x]  bit (y%32) = count[b][y][x] > 0      (32 pixels of a column per word)
// The masks turn fill-hole's per-pixel linear walks (O(W) loads when holes are large: the
// legacy kernel needs 30 ms per 16 frames when the flow converges and most of the frame is a
// hole) into a few word loads plus clz / ffs -- with exactly the same result.
// SIGNED (DepthFlowProjection: count is an accumulated WEIGHT and may be negative or NaN where something landed): the
// bit is count != 0 -- where the reference's walks stop (my_lib_kernel.cu:2210-2225) -- and the fill kernel looks at
// the sign of what it found.
template <bool SIGNED>
__global__ void __launch_bounds__(256) fp_average_mask_kernel(float* __restrict__ out, const float* __restrict__ count,
                                                              unsigned* __restrict__ rowmask, unsigned* __restrict__ colmask,
                                                              int W, int H, int Wt, int Ht, int64_t out_c,
                                                              float* __restrict__ extra) {  // extra: one more plane to divide, or null
    // CTA = 128 x 32 pixels: warp w handles rows w, w+8, w+16, w+24; a lane owns 4 consecutive
    // pixels (128-bit accesses; W % 4 == 0 is a precondition of the fast path)
    __shared__ unsigned rw[32][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 128 + lane * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = warp + 8 * k, y = blockIdx.y * 32 + r;
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x < W && y < H) {
            const int64_t o = (int64_t)y * W + x;
            c = *reinterpret_cast<const float4*>(count + o);
            if (c.x > 0.f || c.y > 0.f || c.z > 0.f || c.w > 0.f) {
                float4 vx = *reinterpret_cast<float4*>(out + o), vy = *reinterpret_cast<float4*>(out + out_c + o);
                if (c.x > 0.f) { vx.x /= c.x; vy.x /= c.x; }
                if (c.y > 0.f) { vx.y /= c.y; vy.y /= c.y; }
                if (c.z > 0.f) { vx.z /= c.z; vy.z /= c.z; }
                if (c.w > 0.f) { vx.w /= c.w; vy.w /= c.w; }
                *reinterpret_cast<float4*>(out + o) = vx;
                *reinterpret_cast<float4*>(out + out_c + o) = vy;
                if (extra) {  // WeightedFlowProjection's weight plane (my_lib_kernel.cu:2653), dense like count
                    float4 e = *reinterpret_cast<float4*>(extra + o);
                    if (c.x > 0.f) e.x /= c.x;
                    if (c.y > 0.f) e.y /= c.y;
                    if (c.z > 0.f) e.z /= c.z;
                    if (c.w > 0.f) e.w /= c.w;
                    *reinterpret_cast<float4*>(extra + o) = e;
                }
            }
        }
        // nibble of this lane -> the 32-pixel word of its 8-lane group (bit = pixel x % 32)
        unsigned nib = (c.x > 0.f ? 1u : 0u) | (c.y > 0.f ? 2u : 0u) | (c.z > 0.f ? 4u : 0u) | (c.w > 0.f ? 8u : 0u);
        if (SIGNED) nib = (c.x != 0.f ? 1u : 0u) | (c.y != 0.f ? 2u : 0u) | (c.z != 0.f ? 4u : 0u) | (c.w != 0.f ? 8u : 0u);
        unsigned word = nib << (4 * (lane & 7));
        word |= __shfl_xor_sync(0xffffffffu, word, 1);
        word |= __shfl_xor_sync(0xffffffffu, word, 2);
        word |= __shfl_xor_sync(0xffffffffu, word, 4);
        if ((lane & 7) == 0) {
            const int wi = blockIdx.x * 4 + (lane >> 3);
            rw[r][lane >> 3] = word;
            if (y < H && wi < Wt) rowmask[(int64_t)y * Wt + wi] = word;
        }
    }
    __syncthreads();
    if (threadIdx.x < 128) {  // column words, stored [y/32][x] so that a warp's loads coalesce
        const int cx = blockIdx.x * 128 + threadIdx.x;
        if (cx < W) {
            unsigned col = 0;
#pragma unroll
            for (int k = 0; k < 32; ++k) col |= ((rw[k][threadIdx.x >> 5] >> (threadIdx.x & 31)) & 1u) << k;
            colmask[(int64_t)blockIdx.y * W + cx] = col;
        }
    }
}

// fill-hole with the masks: identical semantics to flow_projection.cu's fp_fillhole_kernel
// (nearest counted pixel to the left, right and above; never below, my_lib_kernel.cu:1799).
// The word walks fetch four words per step (independent loads) and stop at the first hit.
// SIGNED: the masks mark count != 0; a marked pixel is still a hole when its count is negative, and what a walk finds
// contributes only when it is positive -- the reference's arithmetic on the found counts (my_lib_kernel.cu:2230-2258).
template <bool SIGNED>
__global__ void __launch_bounds__(256) fp_fillhole_mask_kernel(float* __restrict__ out, const unsigned* __restrict__ rowmask,
                                                               const unsigned* __restrict__ colmask, int W, int H, int Wt,
                                                               int Ht, int64_t out_b, int64_t out_c,
                                                               const float* __restrict__ count, int64_t cnt_b) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    const unsigned* rm = rowmask + ((int64_t)b * H + y) * Wt;
    const int wi0 = x >> 5, bit = x & 31;
    const unsigned word = rm[wi0];
    const float* cn = SIGNED ? count + (int64_t)b * cnt_b : nullptr;
    if ((word >> bit) & 1u) {
        if (!SIGNED) return;  // counted pixel: not a hole
        if (!(cn[(int64_t)y * W + x] <= 0.0f)) return;  // positive (or NaN) accumulated weight: not a hole
    }
    int lo = -1, ro = -1, uo = -1;
    {   // left: highest set bit below x
        unsigned m = word & ((1u << bit) - 1u);
        int wi = wi0;
        while (m == 0u && wi > 0) {
            unsigned w4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) w4[k] = (wi - 1 - k >= 0) ? rm[wi - 1 - k] : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (m == 0u && wi > 0) { --wi; m = w4[k]; }
        }
        if (m) lo = wi * 32 + 31 - __clz(m);
    }
    {   // right: lowest set bit above x
        unsigned m = bit == 31 ? 0u : (word & ~((2u << bit) - 1u));
        int wi = wi0;
        while (m == 0u && wi < Wt - 1) {
            unsigned w4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) w4[k] = (wi + 1 + k <= Wt - 1) ? rm[wi + 1 + k] : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (m == 0u && wi < Wt - 1) { ++wi; m = w4[k]; }
        }
        if (m) ro = wi * 32 + __ffs(m) - 1;
    }
    {   // up: highest set bit above y in column x (colmask is [y/32][x])
        const unsigned* cm = colmask + (int64_t)b * Ht * W + x;
        int hi = y >> 5;
        unsigned m = cm[(int64_t)hi * W] & ((1u << (y & 31)) - 1u);
        while (m == 0u && hi > 0) {
            unsigned w4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) w4[k] = (hi - 1 - k >= 0) ? cm[(int64_t)(hi - 1 - k) * W] : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (m == 0u && hi > 0) { --hi; m = w4[k]; }
        }
        if (m) uo = hi * 32 + 31 - __clz(m);
    }
    if (lo < 0 && ro < 0 && uo < 0) return;  // nothing found: stays 0
    float* ox = out + (int64_t)b * out_b;
    float* oy = ox + out_c;
    float sx = 0.f, sy = 0.f, den = 0.f;
    const int64_t row = (int64_t)y * W;
    if (SIGNED) {
        const float lt = lo >= 0 ? cn[row + lo] : 0.f, rt = ro >= 0 ? cn[row + ro] : 0.f;
        const float ut = uo >= 0 ? cn[(int64_t)uo * W + x] : 0.f;
        if (lt + rt + ut <= 0.0f) return;
        if (lt > 0.0f) { sx += ox[row + lo]; sy += oy[row + lo]; den += 1.f; }
        if (rt > 0.0f) { sx += ox[row + ro]; sy += oy[row + ro]; den += 1.f; }
        if (ut > 0.0f) { sx += ox[(int64_t)uo * W + x]; sy += oy[(int64_t)uo * W + x]; den += 1.f; }
        ox[row + x] = sx / den;
        oy[row + x] = sy / den;
        return;
    }
    if (lo >= 0) { sx += ox[row + lo]; sy += oy[row + lo]; den += 1.f; }
    if (ro >= 0) { sx += ox[row + ro]; sy += oy[row + ro]; den += 1.f; }
    if (uo >= 0) { sx += ox[(int64_t)uo * W + x]; sy += oy[(int64_t)uo * W + x]; den += 1.f; }
    ox[row + x] = sx / den;
    oy[row + x] = sy / den;
}

// ------------------------------------------------------------------------------------
// The whole forward as ONE persistent cooperative kernel (MEMC_B200_OVERWRITE calls).
//
// The per-frame sequence above needs 4-5 launches per frame, each far too short (10-30 us) to
// fill the GPU between launch ramp and tail.  Here every CTA keeps pulling work items from ONE
// ordered queue; the items of a "cycle" c are
//
//     average(frame c-2) tiles,  fill-hole(frame c-3) tiles,  splat(frame c) tiles
//
// and an item waits (per-frame completion counters, acquire loads -- no grid-wide barrier) for
//     splat(f)     <- average(f-3) complete   (the accumulator slot f % 3 was handed back zeroed)
//     average(f)   <- splat(f) complete
//     fill-hole(f) <- average(f) complete      (needs the whole frame's occupancy masks)
// Every dependency lies at least a full cycle back in the queue (or behind the cycle's average /
// fill items), so the waits are already satisfied when an item is pulled; all CTAs are
// co-resident (cooperative launch) and dependencies only point backwards in queue order, so the
// oldest unfinished item can always run.
//
// The accumulators live in a three-frame scratch ring (75 MB at 1080p) that stays mostly
// L2-resident: HBM sees the flow once (8 B/px) and the results once (12 B/px).  The flow of the
// NEXT splat tile is prefetched into registers while the current item is processed.
//
// Fill-hole uses two mask levels: the 32-pixel words of the per-frame kernels plus one summary
// bit per word (row: which words of the row are non-empty; column likewise), so a search is a
// handful of loads however large the hole is.
//
// Coherence: everything other CTAs produced inside the launch is read with ld.global.cg (L2);
// the masks (written once, never read before their frame's fill items, 128-byte padded per
// frame) may go through L1.
// ------------------------------------------------------------------------------------
struct FpPipe {
    int B, H, W, fillhole;
    int l2_hints;       // 1: flow loads evict-first, accumulator ring evict-last (createpolicy); 0: no hints
    const float* flowp;
    View flow;
    float* outp;
    float* countp;
    int64_t out_b, out_c, cnt_b;
    float* scratch;     // [3][3][H*W]: sum x, sum y, count of three frames in flight; zero on entry
    unsigned* ctrl;     // [0] queue head, [1 + f] splat tiles of frame f done, [1 + B + f] average tiles done; zero on entry
    unsigned* rowsum;   // [B][rs_stride]  (H x Wt32 words used)   zero on entry
    unsigned* colsum;   // [B][cs_stride]  (Ht32 x W words used)   zero on entry
    unsigned* rowmask;  // [B][rm_stride]  (H x Wt words used)
    unsigned* colmask;  // [B][cm_stride]  (Ht x W words used)
    int64_t rs_stride, cs_stride, rm_stride, cm_stride;  // per-frame strides, multiples of 32 words
    int Wt, Ht, Wt32, Ht32, nS_x, nS_y, nA_x, nA_y;
};

// thread 0: spin until *counter >= target (acquire); a lost producer becomes a launch error
__device__ __forceinline__ void wait_count(const unsigned* counter, unsigned target) {
    for (unsigned spins = 0;; ++spins) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if (v >= target) return;
        __nanosleep(64);
        if (spins > (1u << 27)) {  // several seconds even when time-sliced with another context
            printf("memc_b200: FlowProjection pipeline dependency timed out in block %d\n", blockIdx.x);
            __trap();
        }
    }
}

// average + masks of one 128 x 32 block of frame f (layout of fp_average_mask_kernel); the
// accumulator cells it consumed are handed back zeroed
__device__ __forceinline__ void pipe_average_tile(const FpPipe& p, unsigned (*rw)[4], int idx, int f, uint64_t pol) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bxx = idx % p.nA_x, byy = idx / p.nA_x;
    const int W = p.W, H = p.H;
    const int64_t plane = (int64_t)H * W;
    float* acc = p.scratch + (int64_t)(f % 3) * 3 * plane;
    float* ox = p.outp + (int64_t)f * p.out_b;
    float* oy = ox + p.out_c;
    float* cn = p.countp + (int64_t)f * p.cnt_b;
    unsigned* rowmask = p.rowmask + (int64_t)f * p.rm_stride;
    unsigned* colmask = p.colmask + (int64_t)f * p.cm_stride;
    unsigned* rowsum = p.rowsum + (int64_t)f * p.rs_stride;
    unsigned* colsum = p.colsum + (int64_t)f * p.cs_stride;
    const int x = bxx * 128 + lane * 4;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
        // two rows per round, all six 128-bit loads in flight together (the pass is latency bound)
        float4 c[2], sx[2], sy[2];
        bool in[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int y = byy * 32 + warp + 8 * (2 * kk + h);
            in[h] = x < W && y < H;
            c[h] = sx[h] = sy[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in[h]) {
                const int64_t o = (int64_t)y * W + x;
                c[h] = ldcg4_pol(acc + 2 * plane + o, pol);
                sx[h] = ldcg4_pol(acc + o, pol);
                sy[h] = ldcg4_pol(acc + plane + o, pol);
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = warp + 8 * (2 * kk + h), y = byy * 32 + r;
            const float4 cc = c[h];
            if (in[h]) {
                const int64_t o = (int64_t)y * W + x;
                float4 vx = make_float4(0.f, 0.f, 0.f, 0.f), vy = vx;
                if (cc.x > 0.f || cc.y > 0.f || cc.z > 0.f || cc.w > 0.f) {
                    // my_lib_kernel.cu:1730-1736: divide where count > 0 (nothing was added elsewhere)
                    if (cc.x > 0.f) { vx.x = sx[h].x / cc.x; vy.x = sy[h].x / cc.x; }
                    if (cc.y > 0.f) { vx.y = sx[h].y / cc.y; vy.y = sy[h].y / cc.y; }
                    if (cc.z > 0.f) { vx.z = sx[h].z / cc.z; vy.z = sy[h].z / cc.z; }
                    if (cc.w > 0.f) { vx.w = sx[h].w / cc.w; vy.w = sy[h].w / cc.w; }
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                    stcg4_pol(acc + o, z, pol);
                    stcg4_pol(acc + plane + o, z, pol);
                    stcg4_pol(acc + 2 * plane + o, z, pol);
                }
                __stcg(reinterpret_cast<float4*>(ox + o), vx);  // fill-hole reads neighbours back from L2
                __stcg(reinterpret_cast<float4*>(oy + o), vy);
                __stcs(reinterpret_cast<float4*>(cn + o), cc);
            }
            unsigned nib = (cc.x > 0.f ? 1u : 0u) | (cc.y > 0.f ? 2u : 0u) | (cc.z > 0.f ? 4u : 0u) | (cc.w > 0.f ? 8u : 0u);
            unsigned word = nib << (4 * (lane & 7));
            word |= __shfl_xor_sync(0xffffffffu, word, 1);
            word |= __shfl_xor_sync(0xffffffffu, word, 2);
            word |= __shfl_xor_sync(0xffffffffu, word, 4);
            // summary bits of this row's four words (4 | 32: they share one summary word), on lane 0
            const int wi = bxx * 4 + (lane >> 3);
            unsigned sbit = (word != 0u && wi < p.Wt) ? 1u << (wi & 31) : 0u;
            sbit |= __shfl_xor_sync(0xffffffffu, sbit, 8);
            sbit |= __shfl_xor_sync(0xffffffffu, sbit, 16);
            if ((lane & 7) == 0) {
                rw[r][lane >> 3] = word;
                if (y < H && wi < p.Wt) rowmask[(int64_t)y * p.Wt + wi] = word;
            }
            if (lane == 0 && y < H && sbit) atomicOr(&rowsum[(int64_t)y * p.Wt32 + ((bxx * 4) >> 5)], sbit);
        }
    }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int cx = bxx * 128 + threadIdx.x;
        if (cx < W) {
            unsigned col = 0;
#pragma unroll
            for (int k = 0; k < 32; ++k) col |= ((rw[k][threadIdx.x >> 5] >> (threadIdx.x & 31)) & 1u) << k;
            colmask[(int64_t)byy * W + cx] = col;
            if (col) atomicOr(&colsum[(int64_t)(byy >> 5) * W + cx], 1u << (byy & 31));
        }
    }
}

// highest set bit strictly below position `pos` in a two-level bit set; -1 if none.
// words[i * wstride] = level-1 word i, sums[j * sstride] = summary word j (bit = word non-empty)
__device__ __forceinline__ int find_below(const unsigned* words, int64_t wstride, const unsigned* sums, int64_t sstride,
                                          int pos, unsigned word0) {
    const int wi0 = pos >> 5;
    const unsigned m = word0 & ((1u << (pos & 31)) - 1u);
    if (m) return wi0 * 32 + 31 - __clz(m);
    for (int si = wi0 >> 5; si >= 0; --si) {
        unsigned sm = sums[si * sstride];
        if (si == (wi0 >> 5)) sm &= (1u << (wi0 & 31)) - 1u;  // words strictly below wi0
        if (sm) {
            const int wi = si * 32 + 31 - __clz(sm);
            return wi * 32 + 31 - __clz(words[wi * wstride]);
        }
    }
    return -1;
}
// lowest set bit strictly above `pos` (unit strides); n_sum = number of summary words
__device__ __forceinline__ int find_above(const unsigned* words, const unsigned* sums, int n_sum, int pos, unsigned word0) {
    const int wi0 = pos >> 5, bit = pos & 31;
    const unsigned m = bit == 31 ? 0u : (word0 & ~((2u << bit) - 1u));
    if (m) return wi0 * 32 + __ffs(m) - 1;
    for (int si = wi0 >> 5; si < n_sum; ++si) {
        unsigned sm = sums[si];
        if (si == (wi0 >> 5)) sm = (wi0 & 31) == 31 ? 0u : (sm & ~((2u << (wi0 & 31)) - 1u));  // words strictly above wi0
        if (sm) {
            const int wi = si * 32 + __ffs(sm) - 1;
            return wi * 32 + __ffs(words[wi]) - 1;
        }
    }
    return -1;
}

// fill-hole of one 128 x 32 block of frame f
__device__ __forceinline__ void pipe_fill_tile(const FpPipe& p, unsigned (*rw)[4], int idx, int f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bxx = idx % p.nA_x, byy = idx / p.nA_x;
    const int W = p.W, H = p.H;
    float* ox = p.outp + (int64_t)f * p.out_b;
    float* oy = ox + p.out_c;
    const unsigned* rowmask = p.rowmask + (int64_t)f * p.rm_stride;
    const unsigned* colmask = p.colmask + (int64_t)f * p.cm_stride;
    const unsigned* rowsum = p.rowsum + (int64_t)f * p.rs_stride;
    const unsigned* colsum = p.colsum + (int64_t)f * p.cs_stride;
    // the block's 32 x 4 row words, staged once; most blocks have no hole at all
    int holes = 0;
    if (threadIdx.x < 128) {
        const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
        const int y = byy * 32 + r, wi = bxx * 4 + q;
        unsigned word = 0xffffffffu;
        if (y < H && wi < p.Wt) {
            word = rowmask[(int64_t)y * p.Wt + wi];
            const int valid = W - wi * 32;  // pixels of this word inside the image
            if (valid < 32) word |= ~((1u << valid) - 1u);
        }
        rw[r][q] = word;
        holes = word != 0xffffffffu;
    }
    if (!__syncthreads_or(holes)) return;
#pragma unroll 2
    for (int sub = 0; sub < 16; ++sub) {  // 16 sub-blocks of 32 x 8 (one row per warp)
        const int q = sub & 3, r = (sub >> 2) * 8 + warp;
        const int x = bxx * 128 + q * 32 + lane, y = byy * 32 + r;
        const unsigned word = rw[r][q];
        if ((word >> lane) & 1u) continue;  // counted pixel (or outside the image): not a hole
        const unsigned* rm = rowmask + (int64_t)y * p.Wt;
        const unsigned mword = rm[x >> 5];
        const int lo = find_below(rm, 1, rowsum + (int64_t)y * p.Wt32, 1, x, mword);
        const int ro = find_above(rm, rowsum + (int64_t)y * p.Wt32, p.Wt32, x, mword);
        const unsigned* cm = colmask + x;  // [y/32][x]
        const int uo = find_below(cm, W, colsum + x, W, y, cm[(int64_t)(y >> 5) * W]);
        if (lo < 0 && ro < 0 && uo < 0) continue;  // nothing found: stays 0
        float sx = 0.f, sy = 0.f, den = 0.f;
        const int64_t row = (int64_t)y * W;
        if (lo >= 0) { sx += __ldcg(ox + row + lo); sy += __ldcg(oy + row + lo); den += 1.f; }
        if (ro >= 0) { sx += __ldcg(ox + row + ro); sy += __ldcg(oy + row + ro); den += 1.f; }
        if (uo >= 0) { sx += __ldcg(ox + (int64_t)uo * W + x); sy += __ldcg(oy + (int64_t)uo * W + x); den += 1.f; }
        ox[row + x] = sx / den;
        oy[row + x] = sy / den;
    }
}

// queue position -> work item.  type 0 splat, 1 average, 2 fill-hole, -1 nothing (frame out of
// range), -2 end of queue.  Decoded by thread 0 only (runtime divisions) and published through smem.
struct Item {
    int type, frame, tile, tx;  // tx: splat tile column (tile = row), else unused
};
__device__ __forceinline__ Item decode_item(const FpPipe& p, int nS, int nA, int total, int q) {
    Item it;
    it.frame = it.tile = it.tx = 0;
    if (q >= total) { it.type = -2; return it; }
    const int per_cycle = 2 * nA + nS;
    const int c = q / per_cycle, r = q - c * per_cycle;
    if (r < nA) { it.type = 1; it.frame = c - 2; it.tile = r; }
    else if (r < 2 * nA) { it.type = 2; it.frame = c - 3; it.tile = r - nA; }
    else { it.type = 0; it.frame = c; it.tile = (r - 2 * nA) / p.nS_x; it.tx = (r - 2 * nA) - it.tile * p.nS_x; }
    if (it.frame < 0 || it.frame >= p.B || (it.type == 2 && !p.fillhole)) it.type = -1;
    return it;
}

// completion signal: the release orders the CTA's writes and reductions (observed by this thread
// through the preceding bar.sync) before the counter update
__device__ __forceinline__ void signal_done(unsigned* counter, unsigned n) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(n) : "memory");
}

__global__ void __launch_bounds__(NT, 4) fp_pipeline_kernel(const FpPipe p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    __shared__ Item s_q[3];  // items fetched ahead by thread 0
    unsigned (*rw)[4] = reinterpret_cast<unsigned(*)[4]>(&s.box[0][0][0]);  // 32 x 4 words, aliases the box
    const int tid = threadIdx.x;
    const int nS = p.nS_x * p.nS_y, nA = p.nA_x * p.nA_y;
    const int64_t plane = (int64_t)p.H * p.W;
    const int total = (p.B + 3) * (2 * nA + nS);
    unsigned* done_S = p.ctrl + 1;
    unsigned* done_A = p.ctrl + 1 + p.B;
    int known_S = -1, known_A = -1;  // frames up to here are known complete (uniform across the CTA)
    unsigned pending = 0;            // finished items of the current run not yet signalled
    if (tid == 0) {
        s_q[0] = decode_item(p, nS, nA, total, (int)atomicAdd(p.ctrl, 1u));
        s_q[1] = decode_item(p, nS, nA, total, (int)atomicAdd(p.ctrl, 1u));
        s.ctl[0].reset();
        s.ctl[1].reset();
    }
    int splat_par = 0;  // control set of the next splat tile (uniform across the CTA)
    __syncthreads();
    Item it = s_q[0], in = s_q[1];
    int slot = 2;
    float cfx[PPT], cfy[PPT], nfx[PPT], nfy[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) cfx[k] = cfy[k] = nfx[k] = nfy[k] = 0.f;
    const uint64_t pol_flow = l2_policy(p.l2_hints ? 1 : 0), pol_ring = l2_policy(p.l2_hints ? 2 : 0);
    if (it.type == 0) load_flow<true>(p.flowp, p.flow, it.tx * TW, it.tile * TH, it.frame, p.W, p.H, cfx, cfy, pol_flow);
    while (it.type != -2) {
        int fetched = 0;
        if (tid == 0) fetched = (int)atomicAdd(p.ctrl, 1u);  // the position after next; consumed at the end of this item
        if (in.type == 0) load_flow<true>(p.flowp, p.flow, in.tx * TW, in.tile * TH, in.frame, p.W, p.H, nfx, nfy, pol_flow);
        {   // dependency of this item: normally long satisfied and already known
            const unsigned* dep = nullptr;
            unsigned need = 0;
            if (it.type == 0 && it.frame >= 3 && known_A < it.frame - 3) { dep = done_A + it.frame - 3; need = nA; known_A = it.frame - 3; }
            if (it.type == 1 && known_S < it.frame) { dep = done_S + it.frame; need = nS; known_S = it.frame; }
            if (it.type == 2 && known_A < it.frame) { dep = done_A + it.frame; need = nA; known_A = it.frame; }
            if (dep) {
                if (tid == 0) wait_count(dep, need);
                __syncthreads();
            }
        }
        if (it.type == 0) {
            {
                float* acc = p.scratch + (int64_t)(it.frame % 3) * 3 * plane;
                splat_tile<true>(s, splat_par, cfx, cfy, it.tx * TW, it.tile * TH, acc, acc + plane, acc + 2 * plane, p.W, p.W, p.W,
                                 p.H, pol_ring);
                splat_par ^= 1;
            }
        } else if (it.type == 1) {
            pipe_average_tile(p, rw, it.tile, it.frame, pol_ring);
        } else if (it.type == 2) {
            pipe_fill_tile(p, rw, it.tile, it.frame);
        }
        if (tid == 0) s_q[slot] = decode_item(p, nS, nA, total, fetched);
        __syncthreads();  // s_q[slot] published; this item is done with shared memory and with its global writes
        // completion counters: one signal per run of same-kind items (a CTA's splat tiles of a frame
        // come in a row); a run ends when the next item differs
        if (it.type == 0 || it.type == 1) {
            ++pending;
            if (in.type != it.type || in.frame != it.frame) {
                if (tid == 0) signal_done((it.type == 0 ? done_S : done_A) + it.frame, pending);
                pending = 0;
            }
        }
        it = in;
        in = s_q[slot];
        slot = slot == 2 ? 0 : slot + 1;
#pragma unroll
        for (int k = 0; k < PPT; ++k) { cfx[k] = nfx[k]; cfy[k] = nfy[k]; }
    }
}

size_t pad32(size_t n) { return (n + 31) & ~(size_t)31; }

// OVERWRITE calls: the persistent pipeline.  1 = handled, 0 = not applicable, -1 = error
int fp_forward_pipeline(cudaStream_t stream, const FpArgs& a, int l2_hints) {
    const size_t smem = sizeof(Smem);
    if (!ensure_dynamic_smem(fp_pipeline_kernel, smem)) return 0;
    // device properties and occupancy are looked up once per device (this sits on the launch-bound small-frame path)
    struct DevInfo { int n_sm, per_sm; };
    static DevInfo info[64] = {};
    static std::mutex info_mutex;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    int n_sm, per_sm;
    {
        std::lock_guard<std::mutex> lock(info_mutex);
        if (info[dev].n_sm == 0) {
            int sm = 0, coop = 0, occ = 0;
            cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
            if (!coop || sm <= 0 ||
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fp_pipeline_kernel, NT, smem) != cudaSuccess || occ < 1) {
                cudaGetLastError();
                info[dev] = DevInfo{-1, 0};
            } else {
                info[dev] = DevInfo{sm, occ > 4 ? 4 : occ};
            }
        }
        n_sm = info[dev].n_sm;
        per_sm = info[dev].per_sm;
    }
    if (n_sm <= 0) return 0;
    const int64_t plane = (int64_t)a.H * a.W;
    FpPipe p;
    p.B = a.B; p.H = a.H; p.W = a.W; p.fillhole = a.fillhole; p.l2_hints = l2_hints;
    p.flowp = a.flowp; p.flow = a.flow;
    p.outp = a.outp; p.countp = a.countp;
    p.out_b = a.out.b; p.out_c = a.out.c; p.cnt_b = a.count.b;
    p.Wt = (a.W + 31) / 32; p.Ht = (a.H + 31) / 32;
    p.Wt32 = (p.Wt + 31) / 32; p.Ht32 = (p.Ht + 31) / 32;
    p.nS_x = (a.W + TW - 1) / TW; p.nS_y = (a.H + TH - 1) / TH;
    p.nA_x = (a.W + 127) / 128; p.nA_y = p.Ht;
    p.rs_stride = (int64_t)pad32((size_t)a.H * p.Wt32);
    p.cs_stride = (int64_t)pad32((size_t)p.Ht32 * a.W);
    p.rm_stride = (int64_t)pad32((size_t)a.H * p.Wt);
    p.cm_stride = (int64_t)pad32((size_t)p.Ht * a.W);
    const size_t n_acc = pad32((size_t)(a.B < 3 ? a.B : 3) * 3 * plane), n_ctrl = pad32((size_t)2 * a.B + 1);  // slots f % 3
    const size_t n_sum = (size_t)a.B * (p.rs_stride + p.cs_stride), n_mask = (size_t)a.B * (p.rm_stride + p.cm_stride);
    // one stream-ordered block: [accumulators | ctrl | summaries || masks]; the first three start zeroed
    float* blk = static_cast<float*>(scratch_alloc(stream, (n_acc + n_ctrl + n_sum + n_mask) * 4));
    if (!blk) return 0;
    p.scratch = blk;
    p.ctrl = reinterpret_cast<unsigned*>(blk + n_acc);
    p.rowsum = p.ctrl + n_ctrl;
    p.colsum = p.rowsum + (size_t)a.B * p.rs_stride;
    p.rowmask = p.colsum + (size_t)a.B * p.cs_stride;
    p.colmask = p.rowmask + (size_t)a.B * p.rm_stride;
    int rc = 1;
    if (cudaMemsetAsync(blk, 0, (n_acc + n_ctrl + n_sum) * 4, stream) != cudaSuccess) rc = -1;
    count_launch();
    if (rc == 1) {
        void* args[] = {&p};
        // a context with fewer SMs than the device (MPS thread percentage, green contexts) refuses the cooperative
        // grid: not an error of the op -- hand the call to the frame-by-frame path (rc 0)
        if (cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(fp_pipeline_kernel), dim3(n_sm * per_sm), dim3(NT), args,
                                        smem, stream) != cudaSuccess) {
            cudaGetLastError();
            rc = 0;
        } else {
            count_launch();
            if (check_launch("FlowProjection forward (persistent pipeline)")) rc = -1;
        }
    }
    scratch_free(stream, blk);
    return rc;
}

}  // namespace

static bool fp_fast_layout_ok(const FpArgs& a) {
    if (a.W < SW || a.H < SH || a.W % 4) return false;
    // dense frames only: per-frame memset and the 128-bit averaging pass want contiguous planes
    const int64_t plane = (int64_t)a.H * a.W;
    if (a.out.h != a.W || a.out.c != plane || a.count.h != a.W) return false;
    if ((reinterpret_cast<uintptr_t>(a.outp) & 15u) || (reinterpret_cast<uintptr_t>(a.countp) & 15u)) return false;
    if (a.out.b % 4 || a.count.b % 4) return false;
    return true;
}

static int fp_splat_frame(cudaStream_t stream, const void* ctx, int b) {
    const FpArgs& a = *static_cast<const FpArgs*>(ctx);
    const dim3 grid((a.W + TW - 1) / TW, (a.H + TH - 1) / TH, 1);
    fp_splat_kernel<<<grid, NT, sizeof(Smem), stream>>>(a, b);
    return 0;
}

int fp_forward_fast(cudaStream_t stream, const FpArgs& a, bool overwrite, bool no_zero, int variant) {
    // variant (MEMC_B200_VARIANT field of the flags): 0 production, 1 = frame-by-frame launches even for B >= 3
    if (!fp_fast_layout_ok(a)) return 0;
    // the library produces every element: persistent pipeline (from 3 frames on; below that its phases
    // cannot overlap across frames and the per-frame launches are quicker: 43 vs 58 us at B = 1, 720p)
    if (overwrite && a.B >= 3 && variant != 1) {
        const int r = fp_forward_pipeline(stream, a, variant == 2 ? 0 : 1);  // variant 2: without the L2 eviction hints
        if (r != 0) return r;
    }
    if (!ensure_dynamic_smem(fp_splat_kernel, sizeof(Smem))) return 0;
    return fp_frames_fast(stream, a, overwrite, no_zero, false, fp_splat_frame, &a, nullptr, 0);
}

// frame by frame: [zero fills] -> splat(b) (the caller's kernel) -> average + occupancy masks, then ONE mask-based fill-hole
// launch over the batch.  A frame's count + output planes (25 MB at 1080p) stay L2-resident between its passes.
int fp_frames_fast(cudaStream_t stream, const FpArgs& a, bool overwrite, bool no_zero, bool signed_counts, FpSplatFn splat,
                   const void* ctx, float* extra, int64_t extra_b) {
    if (!fp_fast_layout_ok(a)) return 0;
    if (extra && ((reinterpret_cast<uintptr_t>(extra) & 15u) || extra_b % 4)) return 0;
    const int64_t plane = (int64_t)a.H * a.W;
    // occupancy masks: stream-ordered scratch (1 bit per pixel, twice), freed on the same stream
    const int Wt = (a.W + 31) / 32, Ht = (a.H + 31) / 32;
    const size_t n_row = (size_t)a.B * a.H * Wt, n_col = (size_t)a.B * a.W * Ht;
    unsigned* masks = static_cast<unsigned*>(scratch_alloc(stream, (n_row + n_col) * sizeof(unsigned)));
    if (!masks) return 0;
    unsigned* rowmask = masks;
    unsigned* colmask = masks + n_row;
    const dim3 mgrid((a.W + 127) / 128, Ht, 1);
    int rc = 1;
    for (int b = 0; b < a.B && rc == 1; ++b) {
        float* outb = a.outp + (int64_t)b * a.out.b;
        float* cntb = a.countp + (int64_t)b * a.count.b;
        if (overwrite && !no_zero) {
            if (cudaMemsetAsync(outb, 0, sizeof(float) * 2 * plane, stream) != cudaSuccess) rc = -1;
            if (cudaMemsetAsync(cntb, 0, sizeof(float) * plane, stream) != cudaSuccess) rc = -1;
            count_launch(2);
        }
        if (splat(stream, ctx, b) != 0) rc = -1;
        if (signed_counts)
            fp_average_mask_kernel<true><<<mgrid, 256, 0, stream>>>(outb, cntb, rowmask + (size_t)b * a.H * Wt,
                                                                     colmask + (size_t)b * Ht * a.W, a.W, a.H, Wt, Ht, a.out.c,
                                                                     extra ? extra + (int64_t)b * extra_b : nullptr);
        else
            fp_average_mask_kernel<false><<<mgrid, 256, 0, stream>>>(outb, cntb, rowmask + (size_t)b * a.H * Wt,
                                                                      colmask + (size_t)b * Ht * a.W, a.W, a.H, Wt, Ht, a.out.c,
                                                                      extra ? extra + (int64_t)b * extra_b : nullptr);
        count_launch(2);
        if (check_launch("FlowProjection splat/average (fast)")) rc = -1;
    }
    // fill-hole once over the whole batch (it only needs the masks and the averaged output)
    if (rc == 1 && a.fillhole) {
        const dim3 fgrid(Wt, (a.H + 7) / 8, a.B);
        if (signed_counts)
            fp_fillhole_mask_kernel<true><<<fgrid, 256, 0, stream>>>(a.outp, rowmask, colmask, a.W, a.H, Wt, Ht, a.out.b, a.out.c,
                                                                     a.countp, a.count.b);
        else
            fp_fillhole_mask_kernel<false><<<fgrid, 256, 0, stream>>>(a.outp, rowmask, colmask, a.W, a.H, Wt, Ht, a.out.b, a.out.c,
                                                                      nullptr, 0);
        count_launch();
        if (check_launch("FlowProjection fill-hole (masks)")) rc = -1;
    }
    scratch_free(stream, masks);
    return rc;
}

}  // namespace memc
