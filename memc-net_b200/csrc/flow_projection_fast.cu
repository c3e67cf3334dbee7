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
//   TMA   two reduce-adds (output box, count box) flush the tile: ~1 L2 reduction sector per
//         pixel instead of ~6.6.  Targets outside the box fall back to global atomics.
//
// The frames are processed ONE AT A TIME (memset -> splat -> average -> fill-hole per frame): a
// frame's count+output planes (25 MB at 1080p) then stay L2-resident between the passes instead
// of making three trips to HBM per pass over the whole batch.
#include "flow_projection.cuh"
#include "tma_utils.cuh"
#include <stdlib.h>

namespace memc {

namespace {

constexpr int TW = 64, TH = 8, NT = 256, PPT = TW * TH / NT;  // source tile, 2 pixels / thread
constexpr int SW = 96, SH = 24;                               // target box (pitch 96 = 3*32 words)
constexpr int BOX = SW * SH;

struct __align__(128) Smem {
    float flow[2][TH][TW];  // 4 KB
    int box[3][SH][SW];     // x, y (fixed point) and count: 27 KB; converted to fp32 in place
    uint64_t bar;
    int bb[4];
    unsigned maxbits;
};

__device__ __forceinline__ bool fp_valid(float x2, float y2, int W, int H) {
    return x2 >= 0.0f && y2 >= 0.0f && x2 <= (float)(W - 1) && y2 <= (float)(H - 1);
}

__global__ void __launch_bounds__(NT, 4)
fp_splat_kernel(const __grid_constant__ CUtensorMap m_flow, const __grid_constant__ CUtensorMap m_out,
                const __grid_constant__ CUtensorMap m_count, const FpArgs p, const int b) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw + ((128u - (tma::smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int W = p.W, H = p.H;

    if (tid == 0) {
        tma::mbar_init(&s.bar, 1);
        s.bb[0] = INT_MAX; s.bb[1] = INT_MIN; s.bb[2] = INT_MAX; s.bb[3] = INT_MIN;
        s.maxbits = 0u;
        tma::fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        tma::mbar_expect_tx(&s.bar, sizeof(s.flow));
        tma::load_4d(&s.flow[0][0][0], &m_flow, x0, y0, 0, b, &s.bar);
    }
    {   // zero the box while the flow tile flies
        int4* z = reinterpret_cast<int4*>(&s.box[0][0][0]);
        for (int i = tid; i < 3 * BOX / 4; i += NT) z[i] = make_int4(0, 0, 0, 0);
    }
    tma::mbar_wait(&s.bar, 0, 31);

    // ---- targets of my pixels, tile bounding box, tile max |flow|
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
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    unsigned mb = __float_as_uint(mloc);
#pragma unroll
    for (int o = 16; o; o >>= 1) mb = max(mb, __shfl_xor_sync(0xffffffffu, mb, o));
    if (lane == 0 && mnx <= mxx) {
        atomicMin(&s.bb[0], mnx); atomicMax(&s.bb[1], mxx);
        atomicMin(&s.bb[2], mny); atomicMax(&s.bb[3], mxy);
        if (mb) atomicMax(&s.maxbits, mb);
    }
    __syncthreads();  // also publishes the zeroed box
    if (s.bb[0] > s.bb[1]) return;  // no valid source pixel in this tile (nothing in flight)

    // target cells span [min L, max L + 1] x [min T, max T + 1]; box x origin multiple of 4 (TMA)
    int bx = s.bb[0], by = s.bb[2];
    {
        const int need_w = s.bb[1] - s.bb[0] + 2 + 3, need_h = s.bb[3] - s.bb[2] + 2;
        if (need_w > SW) bx += (need_w - SW) / 2;
        if (need_h > SH) by += (need_h - SH) / 2;
        bx = max(0, min(bx, W - SW)) & ~3;
        by = max(0, min(by, H - SH));
    }
    // fixed-point scale (see header); a non-finite max (NaN flows are invalid, so only Inf-free
    // finite values reach here) cannot occur: valid pixels have finite targets
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

    float* ox = p.outp + (int64_t)b * p.out.b;
    float* oy = ox + p.out.c;
    float* cn = p.countp + (int64_t)b * p.count.b;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        if (!ok[k]) continue;
        const int R = min(L[k] + 1, W - 1), Bm = min(T[k] + 1, H - 1);
        const int qx = __float2int_rn(-fx[k] * scale), qy = __float2int_rn(-fy[k] * scale);
        const int cxs[2] = {L[k], R}, cys[2] = {T[k], Bm};
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int cx = cxs[i], cy = cys[j];
                const int ux = cx - bx, uy = cy - by;
                // a clamped R / Bm repeats a cell (my_lib_kernel.cu:1673-1689): the repeat goes to
                // global memory so that a box cell gets at most one hit per source pixel
                const bool dup = (i == 1 && R == L[k]) || (j == 1 && Bm == T[k]);
                if (!dup && (unsigned)ux < (unsigned)SW && (unsigned)uy < (unsigned)SH) {
                    atomicAdd(&s.box[0][uy][ux], qx);
                    atomicAdd(&s.box[1][uy][ux], qy);
                    atomicAdd(&s.box[2][uy][ux], 1);
                } else {
                    red_add(ox + (int64_t)cy * p.out.h + cx, -fx[k]);
                    red_add(oy + (int64_t)cy * p.out.h + cx, -fy[k]);
                    red_add(cn + (int64_t)cy * p.count.h + cx, 1.0f);
                }
            }
    }
    __syncthreads();
    {   // fixed point -> fp32 in place
        float* bf = reinterpret_cast<float*>(&s.box[0][0][0]);
        const int* bi = &s.box[0][0][0];
        for (int i = tid; i < 2 * BOX; i += NT) bf[i] = (float)bi[i] * inv_scale;
        for (int i = 2 * BOX + tid; i < 3 * BOX; i += NT) bf[i] = (float)bi[i];
    }
    tma::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        tma::reduce_add_4d(&m_out, bx, by, 0, b, &s.box[0][0][0]);
        tma::reduce_add_4d(&m_count, bx, by, 0, b, &s.box[2][0][0]);
        tma::bulk_commit();
        tma::bulk_wait_read_all();
    }
}

// out /= count where count > 0, four pixels per thread (128-bit accesses)
__global__ void __launch_bounds__(256) fp_average4_kernel(float* __restrict__ out, const float* __restrict__ count,
                                                          int64_t out_c, int64_t n4) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n4) return;
    const float4 c = reinterpret_cast<const float4*>(count)[i];
    float4* px = reinterpret_cast<float4*>(out) + i;
    float4* py = reinterpret_cast<float4*>(out + out_c) + i;
    float4 vx = *px, vy = *py;
    if (c.x > 0.f) { vx.x /= c.x; vy.x /= c.x; }
    if (c.y > 0.f) { vx.y /= c.y; vy.y /= c.y; }
    if (c.z > 0.f) { vx.z /= c.z; vy.z /= c.z; }
    if (c.w > 0.f) { vx.w /= c.w; vy.w /= c.w; }
    *px = vx;
    *py = vy;
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
    CUtensorMap m_flow, m_out, m_count;
    if (!tma::make_map_nchw(&m_flow, a.flowp, a.B, 2, a.H, a.W, a.flow.b, a.flow.c, a.flow.h, TW, TH, 2,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
        return 0;
    if (!tma::make_map_nchw(&m_out, a.outp, a.B, 2, a.H, a.W, a.out.b, a.out.c, a.out.h, SW, SH, 2,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE))
        return 0;
    if (!tma::make_map_nchw(&m_count, a.countp, a.B, 1, a.H, a.W, a.count.b, plane, a.count.h, SW, SH, 1,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE))
        return 0;
    static bool configured = false;
    const size_t smem = sizeof(Smem) + 128;
    if (!configured) {
        if (cudaFuncSetAttribute(fp_splat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        configured = true;
    }
    const dim3 grid((a.W + TW - 1) / TW, (a.H + TH - 1) / TH, 1);
    const int64_t n4 = plane / 4;
    for (int b = 0; b < a.B; ++b) {
        float* outb = a.outp + (int64_t)b * a.out.b;
        float* cntb = a.countp + (int64_t)b * a.count.b;
        if (overwrite && !no_zero) {
            if (cudaMemsetAsync(outb, 0, sizeof(float) * 2 * plane, stream) != cudaSuccess) return -1;
            if (cudaMemsetAsync(cntb, 0, sizeof(float) * plane, stream) != cudaSuccess) return -1;
            count_launch(2);
        }
        fp_splat_kernel<<<grid, NT, smem, stream>>>(m_flow, m_out, m_count, a, b);
        fp_average4_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(outb, cntb, a.out.c, n4);
        count_launch(2);
        if (check_launch("FlowProjection splat/average (fast)")) return -1;
        if (a.fillhole && fp_average_fill(stream, a, b, 1, false) != 0) return -1;
    }
    return 1;
}

}  // namespace memc
