// flow_projection_fast.cu -- FlowProjection forward for sm_100a: shared-memory privatised splat.
//
// Same arithmetic as flow_projection.cu (reference my_lib_kernel.cu:1630-1836), different data
// movement.  The legacy splat issues 12 global float atomics per source pixel (ncu: ~6.6 L2
// reduction sectors per pixel, L2-atomic bound; 31 ms per 16 frames when flows converge).  Here a
// CTA owns a TW x TH tile of SOURCE pixels:
//
//   TMA   flow tile (TW,TH,2) -> smem
//   ...   targets p + flow, bounding box of the 2x2 target cells, M = max|flow| over the tile
//   ...   the tile's contributions (-fx, -fy, +1) are accumulated in a SHARED-MEMORY box of
//         3 planes (x, y, count) with native int32 shared atomics: count is an integer anyway,
//         -fx / -fy are accumulated in fixed point with a per-tile power-of-two scale
//         2^e, M * (TW*TH) * 2^e < 2^31 (a cell receives at most one unclamped hit per pixel;
//         border-clamped duplicate hits bypass the box) -- order independent within the tile;
//   ...   the touched part of the box is flushed with 128-bit vector reductions (4 cells each,
//         all-zero vectors skipped).  Targets outside the box fall back to global atomics.
//
// The frames are processed ONE AT A TIME (memset -> splat -> average -> fill-hole per frame): a
// frame's count+output planes (25 MB at 1080p) then stay L2-resident between the passes instead
// of making three trips to HBM per pass over the whole batch.
#include "flow_projection.cuh"
#include "tma_utils.cuh"
#include <stdlib.h>

namespace memc {

namespace {

constexpr int TW = 64, TH = 16, NT = 256, PPT = TW * TH / NT;  // source tile, 4 pixels / thread
constexpr int SW = 96, SH = 32;                                // target box (pitch 96 = 3*32 words)
constexpr int BOX = SW * SH;

struct __align__(128) Smem {
    float flow[2][TH][TW];  // 8 KB
    int box[3][SH][SW];     // x, y (fixed point) and count: 36 KB
    uint64_t bar;
    int bb[4];
    unsigned maxbits;
};

__device__ __forceinline__ bool fp_valid(float x2, float y2, int W, int H) {
    return x2 >= 0.0f && y2 >= 0.0f && x2 <= (float)(W - 1) && y2 <= (float)(H - 1);
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
// Preconditions: a __syncthreads() separates this call from the CTA's previous use of `s`;
// s.bar is an initialised mbarrier whose next phase parity is `phase`.
__device__ __forceinline__ void splat_tile(Smem& s, const CUtensorMap* m_flow, int x0, int y0, int b, unsigned phase,
                                           float* ox, float* oy, float* cn, int64_t out_h, int64_t cnt_h, int W, int H) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s.bb[0] = INT_MAX; s.bb[1] = INT_MIN; s.bb[2] = INT_MAX; s.bb[3] = INT_MIN;
        s.maxbits = 0u;
        tma::fence_proxy_async();  // the flow slot was last read with generic loads
        tma::mbar_expect_tx(&s.bar, sizeof(s.flow));
        tma::load_4d(&s.flow[0][0][0], m_flow, x0, y0, 0, b, &s.bar);
    }
    {   // zero the box while the flow tile flies
        int4* z = reinterpret_cast<int4*>(&s.box[0][0][0]);
        for (int i = tid; i < 3 * BOX / 4; i += NT) z[i] = make_int4(0, 0, 0, 0);
    }
    __syncthreads();  // bounding-box cells initialised, box zeroed
    tma::mbar_wait(&s.bar, phase, 31);

    // ---- targets of my pixels, tile bounding box of the corner cells, tile max |flow|
    float fx[PPT], fy[PPT];
    int L[PPT], T[PPT];
    bool ok[PPT];
    int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
    float mloc = 0.f;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int seg = warp + k * (NT / 32);
        const int yl = seg / (TW / 32), xl = lane + 32 * (seg % (TW / 32));
        const int x = x0 + xl, y = y0 + yl;
        fx[k] = s.flow[0][yl][xl];
        fy[k] = s.flow[1][yl][xl];
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
    if (lane == 0 && mnx <= mxx) {
        atomicMin(&s.bb[0], mnx); atomicMax(&s.bb[1], mxx);
        atomicMin(&s.bb[2], mny); atomicMax(&s.bb[3], mxy);
        if (mb) atomicMax(&s.maxbits, mb);
    }
    __syncthreads();
    if (s.bb[0] > s.bb[1]) return;  // no valid source pixel in this tile (uniform; nothing in flight)

    // corner cells span [min L, max L] x [min T, max T]; box x origin multiple of 4 (vector flush)
    int bx = s.bb[0], by = s.bb[2];
    {
        const int need_w = s.bb[1] - s.bb[0] + 1 + 3, need_h = s.bb[3] - s.bb[2] + 1;
        if (need_w > SW) bx += (need_w - SW) / 2;
        if (need_h > SH) by += (need_h - SH) / 2;
        bx = max(0, min(bx, W - SW)) & ~3;
        by = max(0, min(by, H - SH));
    }
    // fixed-point scale (see header); valid pixels have finite flows, so M is finite
    const float M = __uint_as_float(s.maxbits);
    float scale = 1.0f, inv_scale = 1.0f;
    if (M > 0.f) {
        int ex;
        frexpf(M, &ex);
        constexpr int LOG2_PX = 31 - __builtin_clz(TW * TH - 1) + 1;
        const int e = max(-120, min(31 - ex - LOG2_PX, 120));
        scale = ldexpf(1.0f, e);
        inv_scale = ldexpf(1.0f, -e);
    }

#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        if (!ok[k]) continue;
        const int ux = L[k] - bx, uy = T[k] - by;
        const bool in_box = (unsigned)ux < (unsigned)SW && (unsigned)uy < (unsigned)SH;
        if (in_box) {
            atomicAdd(&s.box[0][uy][ux], __float2int_rn(-fx[k] * scale));
            atomicAdd(&s.box[1][uy][ux], __float2int_rn(-fy[k] * scale));
            atomicAdd(&s.box[2][uy][ux], 1);
        }
        const bool last_col = L[k] == W - 1, last_row = T[k] == H - 1;
        if (__builtin_expect(!in_box || last_col || last_row, 0)) {
            const int R = min(L[k] + 1, W - 1), Bm = min(T[k] + 1, H - 1);
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    // in the box: only the clamped repeats are missing from the 2x2 pattern
                    if (in_box && !((i == 1 && last_col) || (j == 1 && last_row))) continue;
                    const int cx = i ? R : L[k], cy = j ? Bm : T[k];
                    red_add(ox + (int64_t)cy * out_h + cx, -fx[k]);
                    red_add(oy + (int64_t)cy * out_h + cx, -fy[k]);
                    red_add(cn + (int64_t)cy * cnt_h + cx, 1.0f);
                }
        }
    }
    __syncthreads();
    // ---- flush: 2x2 box filter over the touched corner cells, four output cells per 128-bit
    // vector reduction (REDG.E.ADD.F32x4), all-zero vectors skipped.  Output cells beyond the image
    // are dropped (their taps were sent directly above).
    {
        const int ax0 = max(s.bb[0], bx) - bx, ax1 = min(s.bb[1], bx + SW - 1) - bx;
        const int ay0 = max(s.bb[2], by) - by, ay1 = min(s.bb[3], by + SH - 1) - by;
        if (ax0 <= ax1 && ay0 <= ay1) {
            const int cx1 = min(ax1 + 1, W - 1 - bx), cy1 = min(ay1 + 1, H - 1 - by);
            const int v0 = ax0 >> 2, nv = (cx1 >> 2) - v0 + 1, nr = cy1 - ay0 + 1;
            for (int i = tid; i < 3 * nr * nv; i += NT) {
                const int pl = i / (nr * nv), r = i - pl * nr * nv;
                const int uy = ay0 + r / nv, ux = (v0 + r % nv) << 2;  // ux <= SW (a multiple of 4)
                int a[2][5];
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int yy = uy - 1 + rr;
                    const bool row_ok = (unsigned)yy < (unsigned)SH;
                    a[rr][0] = (row_ok && ux > 0) ? s.box[pl][yy][ux - 1] : 0;
                    int4 q = make_int4(0, 0, 0, 0);
                    if (row_ok && ux < SW) q = *reinterpret_cast<const int4*>(&s.box[pl][yy][ux]);
                    a[rr][1] = q.x; a[rr][2] = q.y; a[rr][3] = q.z; a[rr][4] = q.w;
                }
                int q[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) q[k] = a[0][k] + a[0][k + 1] + a[1][k] + a[1][k + 1];
                if ((q[0] | q[1] | q[2] | q[3]) == 0) continue;
                const float sc = pl == 2 ? 1.0f : inv_scale;
                const float4 v = make_float4((float)q[0] * sc, (float)q[1] * sc, (float)q[2] * sc, (float)q[3] * sc);
                float* dst = (pl == 0 ? ox : pl == 1 ? oy : cn) + (int64_t)(by + uy) * (pl == 2 ? cnt_h : out_h) + bx + ux;
                atomicAdd(reinterpret_cast<float4*>(dst), v);
            }
        }
    }
}

__global__ void __launch_bounds__(NT, 4)
fp_splat_kernel(const __grid_constant__ CUtensorMap m_flow, const FpArgs p, const int b) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw + ((128u - (tma::smem_u32(smem_raw) & 127u)) & 127u));
    if (threadIdx.x == 0) {
        tma::mbar_init(&s.bar, 1);
        tma::fence_barrier_init();
    }
    __syncthreads();
    float* ox = p.outp + (int64_t)b * p.out.b;
    splat_tile(s, &m_flow, blockIdx.x * TW, blockIdx.y * TH, b, 0u, ox, ox + p.out.c, p.countp + (int64_t)b * p.count.b,
               p.out.h, p.count.h, p.W, p.H);
}

// ------------------------------------------------------------------------------------
// average + occupancy bit masks.  One CTA = one 128 x 32 pixel block:
//   out /= count where count > 0 (my_lib_kernel.cu:1730-1736), and
//   rowmask[b][y][x/32]  bit (x%32) = count[b][y][x] > 0      (32 pixels of a row per word)
//   colmask[b][y/32][x]  bit (y%32) = count[b][y][x] > 0      (32 pixels of a column per word)
// The masks turn fill-hole's per-pixel linear walks (O(W) loads when holes are large: the
// legacy kernel needs 30 ms per 16 frames when the flow converges and most of the frame is a
// hole) into a few word loads plus clz / ffs -- with exactly the same result.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp_average_mask_kernel(float* __restrict__ out, const float* __restrict__ count,
                                                              unsigned* __restrict__ rowmask, unsigned* __restrict__ colmask,
                                                              int W, int H, int Wt, int Ht, int64_t out_c) {
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
            }
        }
        // nibble of this lane -> the 32-pixel word of its 8-lane group (bit = pixel x % 32)
        unsigned nib = (c.x > 0.f ? 1u : 0u) | (c.y > 0.f ? 2u : 0u) | (c.z > 0.f ? 4u : 0u) | (c.w > 0.f ? 8u : 0u);
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
__global__ void __launch_bounds__(256) fp_fillhole_mask_kernel(float* __restrict__ out, const unsigned* __restrict__ rowmask,
                                                               const unsigned* __restrict__ colmask, int W, int H, int Wt,
                                                               int Ht, int64_t out_b, int64_t out_c) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    const unsigned* rm = rowmask + ((int64_t)b * H + y) * Wt;
    const int wi0 = x >> 5, bit = x & 31;
    const unsigned word = rm[wi0];
    if ((word >> bit) & 1u) return;  // counted pixel: not a hole
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
    if (lo >= 0) { sx += ox[row + lo]; sy += oy[row + lo]; den += 1.f; }
    if (ro >= 0) { sx += ox[row + ro]; sy += oy[row + ro]; den += 1.f; }
    if (uo >= 0) { sx += ox[(int64_t)uo * W + x]; sy += oy[(int64_t)uo * W + x]; den += 1.f; }
    ox[row + x] = sx / den;
    oy[row + x] = sy / den;
}

// Library-owned stream-ordered memory pool for the occupancy masks (the only scratch memory the
// library ever allocates).  A private pool with a high release threshold keeps the blocks
// cached across calls; the device's default pool would hand them back to the OS at every
// synchronisation (measured: 2 ms per call).  One pool per device, created on first use.
cudaMemPool_t scratch_pool() {
    static cudaMemPool_t pools[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        pools[dev] = pool;
    }
    return pools[dev];
}

}  // namespace

int fp_forward_fast(cudaStream_t stream, const FpArgs& a, bool overwrite, bool no_zero) {
    if (const char* e = getenv("MEMC_TMA_DBG")) if (atoi(e) & 32) return 0;  // development: generic path
    if (a.W < SW || a.H < SH || a.W % 4) return 0;
    // dense frames only: per-frame memset and the 128-bit averaging pass want contiguous planes
    const int64_t plane = (int64_t)a.H * a.W;
    if (a.out.h != a.W || a.out.c != plane || a.count.h != a.W) return 0;
    if ((reinterpret_cast<uintptr_t>(a.outp) & 15u) || (reinterpret_cast<uintptr_t>(a.countp) & 15u)) return 0;
    if (a.out.b % 4 || a.count.b % 4) return 0;
    CUtensorMap m_flow;
    if (!tma::make_map_nchw(&m_flow, a.flowp, a.B, 2, a.H, a.W, a.flow.b, a.flow.c, a.flow.h, TW, TH, 2,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
        return 0;
    const size_t smem = sizeof(Smem) + 128;
    if (!ensure_dynamic_smem(fp_splat_kernel, smem)) return 0;
    const dim3 grid((a.W + TW - 1) / TW, (a.H + TH - 1) / TH, 1);
    // occupancy masks: stream-ordered scratch (1 bit per pixel, twice), freed on the same stream
    const int Wt = (a.W + 31) / 32, Ht = (a.H + 31) / 32;
    const size_t n_row = (size_t)a.B * a.H * Wt, n_col = (size_t)a.B * a.W * Ht;
    unsigned* masks = nullptr;
    cudaMemPool_t pool = scratch_pool();
    if (!pool || cudaMallocFromPoolAsync(reinterpret_cast<void**>(&masks), (n_row + n_col) * sizeof(unsigned), pool,
                                         stream) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
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
        fp_splat_kernel<<<grid, NT, smem, stream>>>(m_flow, a, b);
        fp_average_mask_kernel<<<mgrid, 256, 0, stream>>>(outb, cntb, rowmask + (size_t)b * a.H * Wt,
                                                           colmask + (size_t)b * Ht * a.W, a.W, a.H, Wt, Ht, a.out.c);
        count_launch(2);
        if (check_launch("FlowProjection splat/average (fast)")) rc = -1;
    }
    // fill-hole once over the whole batch (it only needs the masks and the averaged output)
    if (rc == 1 && a.fillhole) {
        const dim3 fgrid(Wt, (a.H + 7) / 8, a.B);
        fp_fillhole_mask_kernel<<<fgrid, 256, 0, stream>>>(a.outp, rowmask, colmask, a.W, a.H, Wt, Ht, a.out.b, a.out.c);
        count_launch();
        if (check_launch("FlowProjection fill-hole (masks)")) rc = -1;
    }
    cudaFreeAsync(masks, stream);
    return rc;
}

}  // namespace memc
