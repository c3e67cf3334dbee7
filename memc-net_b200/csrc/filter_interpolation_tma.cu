// filter_interpolation_tma.cu -- fs = 4 fast path of FilterInterpolation for sm_100a.
//
// Same arithmetic as the generic kernels (filter_interpolation.cu; reference
// my_lib_kernel.cu:1087-1518), different data movement.  The generic kernel is bound by
// memory-level parallelism: every byte in flight is pinned to a register of a resident thread
// (Little's law: ~64 KB must be in flight per SM to saturate HBM3e) and by L1 wavefronts (a
// warp-wide gather of 32 neighbouring pixels straddles two 128-byte lines).  Here one elected
// thread per CTA hands the tile to the TMA engine and the taps are served from shared memory:
//
//   tile   = TW x TH output pixels of one frame (warp-wide row segments of 32 pixels)
//   TMA 1  flow   box (TW,TH,2)   -> smem  } issued first, no dependence on anything
//   TMA 2  filter box (TW,TH,16)  -> smem  } (forward; backward reads the filter into registers)
//   ...    threads read the flow, compute the integer target of every pixel and reduce the
//          tile's bounding box of source windows (data dependent!)
//   TMA 3  image  box (SW,SH,C) at that origin -> smem   (clamped taps stay inside the image;
//          whatever the box misses is fetched with plain loads -- always correct)
//   ...    filter taps and image taps come from shared memory (row pitch SW = k*32 words, so
//          the bank of a tap depends on its column only); results leave through coalesced STG.
//   backward additionally accumulates gradinput1 in a shared-memory box congruent with the
//   image box and flushes it with ONE TMA reduce-add per tile (UTMAREDG, performed by the L2).
//
// Several CTAs share an SM so that one computes while the others' loads are in flight.
// Layout preconditions (else the caller falls back to the generic kernel): fs == 4, C <= 4,
// W % 4 == 0, 16-byte aligned bases and strides (TMA), H >= SH, W >= SW.
//
// Measured TMA constraint (B200, driver 580): the innermost tile coordinate must be a multiple
// of 16 bytes (an unaligned c0 raises "illegal instruction"), hence the box origin x is rounded
// down to a multiple of 4 pixels.
#include "filter_interpolation.cuh"
#include "tma_utils.cuh"

namespace memc {

namespace {

constexpr int CB = 4;  // max channels staged per box

// TW x TH output tile, SW x SH image box, NT threads, MINB resident CTAs per SM (launch bound)
// PATCH: a warp instruction covers an 8 x 4 pixel patch instead of a 32-pixel row segment (fewer
// bank conflicts in the tap gather when SW % 32 == 8, profiles/r01_bank_conflict_model.md)
template <int TW_, int TH_, int SW_, int SH_, int NT_, int MINB_, bool PATCH_ = false>
struct Cfg {
    static constexpr int TW = TW_, TH = TH_, SW = SW_, SH = SH_, NT = NT_, MINB = MINB_;
    static constexpr bool PATCH = PATCH_;
    static constexpr int SEGS = TW / 32;      // 32-pixel row segments per tile row
    static constexpr int PPT = TW * TH / NT;  // pixels per thread
    static_assert(TW % 32 == 0 && (PATCH_ ? SW % 4 == 0 && TH % 4 == 0 : SW % 32 == 0) && (TW * TH) % NT == 0 && NT % 32 == 0,
                  "bad tile config");
};

__device__ __forceinline__ int warp_min(int v) { return __reduce_min_sync(0xffffffffu, v); }  // REDUX.MIN.S32
__device__ __forceinline__ int warp_max(int v) { return __reduce_max_sync(0xffffffffu, v); }

// shared-memory carve-up (byte offsets from a 128-byte aligned base)
struct Layout {
    int off_flow, off_a, off_bar, off_img, off_acc, total;
};
template <class K>
__host__ __device__ constexpr Layout make_layout(int C, int planes_a, bool with_acc) {
    Layout l{};
    l.off_a = 0;  // filter (fwd) / gradoutput (bwd) tile
    l.off_flow = planes_a * K::TH * K::TW * 4;
    l.off_bar = l.off_flow + 2 * K::TH * K::TW * 4;
    l.off_img = (l.off_bar + 64 + 127) & ~127;
    l.off_acc = l.off_img + C * K::SH * K::SW * 4;
    l.total = with_acc ? l.off_acc + C * K::SH * K::SW * 4 : l.off_acc;
    return l;
}

// pixel k of this thread inside the tile: a warp always owns a 32-pixel row segment
template <class K>
__device__ __forceinline__ void tile_pixel(int k, int lane, int warp, int& xl, int& yl) {
    const int seg = warp + k * (K::NT / 32);
    if (K::PATCH) {  // patch `seg` of the tile: TW/8 patches per patch row
        yl = 4 * (seg / (K::TW / 8)) + (lane >> 3);
        xl = 8 * (seg % (K::TW / 8)) + (lane & 7);
    } else {
        yl = seg / K::SEGS;
        xl = lane + 32 * (seg % K::SEGS);
    }
}

// bounding box of the source windows over the tile's valid pixels -> box origin (bx, by)
template <class K>
__device__ __forceinline__ bool tile_box(const float* s_flow, int* s_bb, int x0, int y0, int W, int H, int lane,
                                         int warp, int& bx, int& by) {
    int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
#pragma unroll
    for (int k = 0; k < K::PPT; ++k) {
        int xl, yl;
        tile_pixel<K>(k, lane, warp, xl, yl);
        const FiGeom g = fi_geometry(x0 + xl, y0 + yl, W, H, s_flow[yl * K::TW + xl], s_flow[(K::TH + yl) * K::TW + xl]);
        if (g.valid && x0 + xl < W && y0 + yl < H) {
            mnx = min(mnx, g.ix); mxx = max(mxx, g.ix);
            mny = min(mny, g.iy); mxy = max(mxy, g.iy);
        }
    }
    mnx = warp_min(mnx); mxx = warp_max(mxx); mny = warp_min(mny); mxy = warp_max(mxy);
    if (lane == 0 && mnx <= mxx) {
        atomicMin(&s_bb[0], mnx); atomicMax(&s_bb[1], mxx);
        atomicMin(&s_bb[2], mny); atomicMax(&s_bb[3], mxy);
    }
    __syncthreads();
    const bool any_valid = s_bb[0] <= s_bb[1];
    if (!any_valid) {  // nothing to stage: any legal origin will do
        bx = 0;
        by = 0;
        return false;
    }
    // windows span [min ix - 1, max ix + 2]; taps are clamped into the image, so the origin is
    // clamped too; a span wider than the box centres the box; x is rounded down to 4 (TMA)
    bx = s_bb[0] - 1;
    by = s_bb[2] - 1;
    const int need_w = s_bb[1] - s_bb[0] + 4 + 3, need_h = s_bb[3] - s_bb[2] + 4;
    if (need_w > K::SW) bx += (need_w - K::SW) / 2;
    if (need_h > K::SH) by += (need_h - K::SH) / 2;
    bx = max(0, min(bx, W - K::SW)) & ~3;  // W >= SW and W % 4 == 0 are launch preconditions
    by = max(0, min(by, H - K::SH));       // H >= SH is a launch precondition
    return any_valid;
}

// ====================================================================================
// forward
// ====================================================================================
// A pixel whose window touches the image border or leaves the staged box (a few lanes of ~5 % of
// the warps on the benchmark field): per-tap clamping, box or global source per tap.  Run AFTER the
// fast pixels of the thread (nothing of the fast path is live any more, so it does not add to the
// register budget), one window row per iteration with the channel loop innermost: ~400 warp
// instructions instead of ~2000 for a tap-by-tap loop per channel.
template <int C, class K>
__device__ __forceinline__ void fwd_slow_pixel(const FiArgs& p, const float* wcol, const int wstride, const float* s_img,
                                               const float* in1b, int L, int T, int bx, int by, float a, float bt,
                                               float (&res)[C]) {
    constexpr int SW = K::SW, SH = K::SH;
    const int W = p.W, H = p.H;
    float q[C][4];
#pragma unroll
    for (int c = 0; c < C; ++c) q[c][0] = q[c][1] = q[c][2] = q[c][3] = 0.f;
    // two window rows (8 taps x C loads, box or L2) are in flight at a time: two memory round trips
    // per slow pixel; the quadrant of a tap is static
#pragma unroll
    for (int jp = 0; jp < 2; ++jp) {
        float v[2][4][C], wt[2][4];
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const int j = 2 * jp + jj;
            const int cy = clampi(T + j, 0, H - 1);
            const int uy = cy - by;
            const bool row_in = (unsigned)uy < (unsigned)SH;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int cx = clampi(L + i, 0, W - 1);
                const int ux = cx - bx;
                const bool in_box = row_in && (unsigned)ux < (unsigned)SW;
                wt[jj][i] = wcol[(j * 4 + i) * wstride];
                const float* gsrc = in1b + (int64_t)cy * p.in1.h + cx;
                const float* ssrc = s_img + uy * SW + ux;
#pragma unroll
                for (int c = 0; c < C; ++c) v[jj][i][c] = in_box ? ssrc[c * SH * SW] : __ldg(gsrc + c * p.in1.c);
            }
        }
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < C; ++c)  // same fmaf chain per (channel, quadrant) as the fast path
                    q[c][2 * jp + (i >> 1)] = fmaf(v[jj][i][c], wt[jj][i], q[c][2 * jp + (i >> 1)]);
    }
    const float wTL = (1.0f - a) * (1.0f - bt), wTR = a * (1.0f - bt);
    const float wBL = (1.0f - a) * bt, wBR = a * bt;
#pragma unroll
    for (int c = 0; c < C; ++c) res[c] = wTL * q[c][0] + wTR * q[c][1] + wBL * q[c][2] + wBR * q[c][3];
}

// all pixels of one staged tile: shared by the one-tile-per-CTA and the persistent kernels
template <int C, class K>
__device__ __forceinline__ void fwd_compute_tile(const FiArgs& p, const float* s_filt, const float* s_flow,
                                                 const float* s_img, int x0, int y0, int b, int bx, int by, int lane,
                                                 int warp) {
    constexpr int TW = K::TW, TH = K::TH, SW = K::SW, SH = K::SH;
    const int W = p.W, H = p.H;
    const float* in1b = p.in1p + b * p.in1.b;
    unsigned slow = 0;  // bit k: pixel k of this thread needs the clamped path
#pragma unroll
    for (int k = 0; k < K::PPT; ++k) {
        int xl, yl;
        tile_pixel<K>(k, lane, warp, xl, yl);
        const int x = x0 + xl, y = y0 + yl;
        if (x >= W || y >= H) continue;
        float* outp = p.outp + b * p.out.b + (int64_t)y * p.out.h + x;
        // geometry is recomputed from the staged flow (cheaper than keeping it live in registers)
        const FiGeom g = fi_geometry(x, y, W, H, s_flow[yl * TW + xl], s_flow[(TH + yl) * TW + xl]);
        if (!g.valid) {  // my_lib_kernel.cu:1209-1213: copy the input pixel
#pragma unroll
            for (int c = 0; c < C; ++c)
                stg_stream(outp + c * p.out.c, __ldg(in1b + c * p.in1.c + (int64_t)y * p.in1.h + x));
            continue;
        }
        const int lx = g.ix - 1 - bx, ly = g.iy - 1 - by;
        // the box lies inside the image (0 <= bx <= W-SW, 0 <= by <= H-SH), so a window inside the
        // box is also inside the image: no per-tap clamping needed
        if (!(((unsigned)lx <= (unsigned)(SW - 4)) && ((unsigned)ly <= (unsigned)(SH - 4)))) {
            slow |= 1u << k;
            continue;
        }
        const float a = g.alpha, bt = g.beta;
        const float wTL = (1.0f - a) * (1.0f - bt), wTR = a * (1.0f - bt);
        const float wBL = (1.0f - a) * bt, wBR = a * bt;
        const float* wcol = s_filt + yl * TW + xl;  // plane stride TH*TW
        float wg[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) wg[t] = wcol[t * TH * TW];
        const float* base = s_img + ly * SW + lx;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    q[(j >> 1) * 2 + (i >> 1)] =
                        fmaf(base[c * SH * SW + j * SW + i], wg[j * 4 + i], q[(j >> 1) * 2 + (i >> 1)]);
            stg_stream(outp + c * p.out.c, wTL * q[0] + wTR * q[1] + wBL * q[2] + wBR * q[3]);
        }
    }
    if (__builtin_expect(slow != 0, 0)) {
#pragma unroll 1
        for (int k = 0; k < K::PPT; ++k) {
            if (!((slow >> k) & 1u)) continue;
            int xl, yl;
            tile_pixel<K>(k, lane, warp, xl, yl);
            const int x = x0 + xl, y = y0 + yl;
            const FiGeom g = fi_geometry(x, y, W, H, s_flow[yl * TW + xl], s_flow[(TH + yl) * TW + xl]);
            float res[C];
            fwd_slow_pixel<C, K>(p, s_filt + yl * TW + xl, TH * TW, s_img, in1b, g.ix - 1, g.iy - 1, bx, by, g.alpha,
                                 g.beta, res);
            float* outp = p.outp + b * p.out.b + (int64_t)y * p.out.h + x;
#pragma unroll
            for (int c = 0; c < C; ++c) stg_stream(outp + c * p.out.c, res[c]);
        }
    }
}

template <int C, class K>
__global__ void __launch_bounds__(K::NT, K::MINB)
fi_fwd_tma_kernel(const __grid_constant__ CUtensorMap m_flow, const __grid_constant__ CUtensorMap m_filt,
                  const __grid_constant__ CUtensorMap m_img, const __grid_constant__ FiArgs p) {
    constexpr int TW = K::TW, TH = K::TH, SW = K::SW, SH = K::SH;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((128u - (tma::smem_u32(smem_raw) & 127u)) & 127u);  // TMA: 128-byte boxes
    constexpr Layout lay = make_layout<K>(C, 16, false);
    const float* s_filt = reinterpret_cast<const float*>(sm + lay.off_a);     // [16][TH][TW]
    const float* s_flow = reinterpret_cast<const float*>(sm + lay.off_flow);  // [2][TH][TW]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + lay.off_bar);           // 0 flow, 1 filter, 2 image
    int* s_bb = reinterpret_cast<int*>(bars + 3);
    const float* s_img = reinterpret_cast<const float*>(sm + lay.off_img);    // [C][SH][SW]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
    const int W = p.W, H = p.H;

    if (tid == 0) {
        tma::mbar_init(&bars[0], 1);
        tma::mbar_init(&bars[1], 1);
        tma::mbar_init(&bars[2], 1);
        s_bb[0] = INT_MAX; s_bb[1] = INT_MIN; s_bb[2] = INT_MAX; s_bb[3] = INT_MIN;
        tma::fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        tma::mbar_expect_tx(&bars[0], 2 * TH * TW * 4);
        tma::load_4d(sm + lay.off_flow, &m_flow, x0, y0, 0, b, &bars[0]);
        tma::mbar_expect_tx(&bars[1], 16 * TH * TW * 4);
        tma::load_4d(sm + lay.off_a, &m_filt, x0, y0, 0, b, &bars[1]);
    }

    tma::mbar_wait(&bars[0], 0, 1);
    int bx, by;
    const bool any_valid = tile_box<K>(s_flow, s_bb, x0, y0, W, H, lane, warp, bx, by);
    if (tid == 0 && any_valid) {
        tma::mbar_expect_tx(&bars[2], C * SH * SW * 4);
        tma::load_4d(sm + lay.off_img, &m_img, bx, by, 0, b, &bars[2]);
    }
    tma::mbar_wait(&bars[1], 0, 2);
    if (any_valid) tma::mbar_wait(&bars[2], 0, 3);

    fwd_compute_tile<C, K>(p, s_filt, s_flow, s_img, x0, y0, b, bx, by, lane, warp);
}

bool make_maps(const FiArgs& a, bool bwd, int TW, int TH, int SW, int SH, CUtensorMap* m);  // m[5]

// ------------------------------------------------------------------------------------
// forward for C > 4 (e.g. the 64-channel context features MEMC_Net_star warps with the same
// flow / filter, networks/MEMC_Net_star.py:280-285): flow tile, filter tile and bounding box are
// set up ONCE per tile, then the image is streamed through a two-buffer ring of CBK-channel
// boxes while the 16 filter taps and the geometry of each pixel stay in registers
// (the legacy kernel re-reads the 16 filter planes for every one of the 64 channels).
// ------------------------------------------------------------------------------------
constexpr int CBK = 4;  // channels per streamed box

template <class K>
__host__ __device__ constexpr int chunked_smem() {
    return 16 * K::TH * K::TW * 4 + 2 * K::TH * K::TW * 4 + 2 * (CBK * K::SH * K::SW * 4) + 256;
}

template <class K>
__global__ void __launch_bounds__(K::NT, K::MINB)
fi_fwd_tma_chunked_kernel(const __grid_constant__ CUtensorMap m_flow, const __grid_constant__ CUtensorMap m_filt,
                          const __grid_constant__ CUtensorMap m_img, const __grid_constant__ FiArgs p) {
    constexpr int TW = K::TW, TH = K::TH, SW = K::SW, SH = K::SH, PPT = K::PPT;
    constexpr int FILT_B = 16 * TH * TW * 4, FLOW_B = 2 * TH * TW * 4, IMG_B = CBK * SH * SW * 4;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((128u - (tma::smem_u32(smem_raw) & 127u)) & 127u);
    const float* s_filt = reinterpret_cast<const float*>(sm);
    const float* s_flow = reinterpret_cast<const float*>(sm + FILT_B);
    unsigned char* const sm_img0 = sm + FILT_B + FLOW_B;                        // 2 buffers, IMG_B apart
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm_img0 + 2 * IMG_B);          // 0 flow, 1 filter, 2/3 image
    int* s_bb = reinterpret_cast<int*>(bars + 4);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
    const int W = p.W, H = p.H, C = p.C;
    const int nchunk = (C + CBK - 1) / CBK;

    if (tid == 0) {
        for (int k = 0; k < 4; ++k) tma::mbar_init(&bars[k], 1);
        s_bb[0] = INT_MAX; s_bb[1] = INT_MIN; s_bb[2] = INT_MAX; s_bb[3] = INT_MIN;
        tma::fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        tma::mbar_expect_tx(&bars[0], FLOW_B);
        tma::load_4d(sm + FILT_B, &m_flow, x0, y0, 0, b, &bars[0]);
        tma::mbar_expect_tx(&bars[1], FILT_B);
        tma::load_4d(sm, &m_filt, x0, y0, 0, b, &bars[1]);
    }
    tma::mbar_wait(&bars[0], 0, 51);
    int bx, by;
    tile_box<K>(s_flow, s_bb, x0, y0, W, H, lane, warp, bx, by);
    if (tid == 0) {
        for (int ch = 0; ch < 2 && ch < nchunk; ++ch) {
            tma::mbar_expect_tx(&bars[2 + ch], IMG_B);
            tma::load_4d(sm_img0 + ch * IMG_B, &m_img, bx, by, ch * CBK, b, &bars[2 + ch]);
        }
    }
    tma::mbar_wait(&bars[1], 0, 52);

    // ---- per-pixel state kept in registers across the channel chunks
    float wg[PPT][16], wq[PPT][4];
    int off[PPT];      // >= 0: tap (0,0) offset inside the box (fast path); -1: slow path; -2: invalid; -3: outside
    int Lx[PPT], Ty[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        int xl, yl;
        tile_pixel<K>(k, lane, warp, xl, yl);
        const int x = x0 + xl, y = y0 + yl;
        off[k] = -3;
        Lx[k] = Ty[k] = 0;
        if (x >= W || y >= H) continue;
        const FiGeom g = fi_geometry(x, y, W, H, s_flow[yl * TW + xl], s_flow[(TH + yl) * TW + xl]);
        if (!g.valid) { off[k] = -2; continue; }
#pragma unroll
        for (int t = 0; t < 16; ++t) wg[k][t] = s_filt[(t * TH + yl) * TW + xl];
        const float a = g.alpha, bt = g.beta;
        wq[k][0] = (1.0f - a) * (1.0f - bt); wq[k][1] = a * (1.0f - bt);
        wq[k][2] = (1.0f - a) * bt;          wq[k][3] = a * bt;
        Lx[k] = g.ix - 1;
        Ty[k] = g.iy - 1;
        const int lx = Lx[k] - bx, ly = Ty[k] - by;
        const bool fast = (Lx[k] >= 0) && (Lx[k] + 3 <= W - 1) && (Ty[k] >= 0) && (Ty[k] + 3 <= H - 1) &&
                          (lx >= 0) && (lx + 3 < SW) && (ly >= 0) && (ly + 3 < SH);
        off[k] = fast ? ly * SW + lx : -1;
    }

    const float* in1b = p.in1p + b * p.in1.b;
    for (int ch = 0; ch < nchunk; ++ch) {
        const float* s_img = reinterpret_cast<const float*>(sm_img0 + (ch & 1) * IMG_B);
        tma::mbar_wait(&bars[2 + (ch & 1)], (ch >> 1) & 1, 53);
        const int c0 = ch * CBK;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            if (off[k] == -3) continue;
            int xl, yl;
            tile_pixel<K>(k, lane, warp, xl, yl);
            const int x = x0 + xl, y = y0 + yl;
            float* outp = p.outp + b * p.out.b + (int64_t)y * p.out.h + x;
#pragma unroll
            for (int cc = 0; cc < CBK; ++cc) {
                const int c = c0 + cc;
                if (c >= C) break;
                float v;
                if (off[k] == -2) {  // my_lib_kernel.cu:1209-1213: copy the input pixel
                    v = __ldg(in1b + c * p.in1.c + (int64_t)y * p.in1.h + x);
                } else if (off[k] >= 0) {
                    const float* base = s_img + cc * SH * SW + off[k];
                    float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            q[(j >> 1) * 2 + (i >> 1)] = fmaf(base[j * SW + i], wg[k][j * 4 + i], q[(j >> 1) * 2 + (i >> 1)]);
                    v = wq[k][0] * q[0] + wq[k][1] * q[1] + wq[k][2] * q[2] + wq[k][3] * q[3];
                } else {
                    const float* img = in1b + c * p.in1.c;
                    float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int cx = clampi(Lx[k] + i, 0, W - 1), cy = clampi(Ty[k] + j, 0, H - 1);
                            const int ux = cx - bx, uy = cy - by;
                            const bool in_box = (unsigned)ux < (unsigned)SW && (unsigned)uy < (unsigned)SH;
                            const float t = in_box ? s_img[cc * SH * SW + uy * SW + ux] : __ldg(img + (int64_t)cy * p.in1.h + cx);
                            q[(j >> 1) * 2 + (i >> 1)] = fmaf(t, wg[k][j * 4 + i], q[(j >> 1) * 2 + (i >> 1)]);
                        }
                    v = wq[k][0] * q[0] + wq[k][1] * q[1] + wq[k][2] * q[2] + wq[k][3] * q[3];
                }
                stg_stream(outp + c * p.out.c, v);
            }
        }
        __syncthreads();  // everybody is done with buffer ch&1
        if (tid == 0 && ch + 2 < nchunk) {
            tma::fence_proxy_async();
            tma::mbar_expect_tx(&bars[2 + (ch & 1)], IMG_B);
            tma::load_4d(sm_img0 + (ch & 1) * IMG_B, &m_img, bx, by, (ch + 2) * CBK, b, &bars[2 + (ch & 1)]);
        }
    }
}

template <class K>
int launch_fwd_chunked(cudaStream_t stream, const FiArgs& a) {
    if (a.W < K::SW || a.H < K::SH) return 0;
    CUtensorMap m[5];
    FiArgs a4 = a;
    if (!tma::make_map_nchw(&m[0], a.flowp, a.B, 2, a.H, a.W, a.flow.b, a.flow.c, a.flow.h, K::TW, K::TH, 2,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B) ||
        !tma::make_map_nchw(&m[1], a.filtp, a.B, 16, a.H, a.W, a.filt.b, a.filt.c, a.filt.h, K::TW, K::TH, 16,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B) ||
        !tma::make_map_nchw(&m[2], a.in1p, a.B, a.C, a.H, a.W, a.in1.b, a.in1.c, a.in1.h, K::SW, K::SH, CBK,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
        return 0;
    constexpr size_t smem = (size_t)chunked_smem<K>() + 128;
    if (!ensure_dynamic_smem(fi_fwd_tma_chunked_kernel<K>, smem)) return 0;
    dim3 grid((a.W + K::TW - 1) / K::TW, (a.H + K::TH - 1) / K::TH, a.B);
    fi_fwd_tma_chunked_kernel<K><<<grid, K::NT, smem, stream>>>(m[0], m[1], m[2], a4);
    count_launch();
    return check_launch("FilterInterpolation forward (TMA, channel-chunked)") == 0 ? 1 : -1;
}

// ------------------------------------------------------------------------------------
// forward, "patch" variant: a warp works on 8x4-pixel patches instead of 32x1 row segments and the
// image box has a row pitch of 72 words; the modelled bank-conflict factor of the 16-tap gather
// drops from 2.40 to 1.96 on the benchmark field (tools/bank_model.py).  Flow and filter tiles are
// staged as four 8-pixel-wide strips ([strip][plane][TH][8], one TMA each) so that a patch reads
// 32 consecutive words.  With STAGE_OUT the results of a strip go back through shared memory (the
// warp's own, by then dead, filter strip) and one TMA store per warp instead of 4-row partial
// stores.  Tile 32 x 8, 128 threads; warp w owns strip w.
// ------------------------------------------------------------------------------------
template <int SW_, int SH_>
struct PatchCfg {
    static constexpr int TW = 32, TH = 8, SW = SW_, SH = SH_;
};

// shared-memory carve-up of the patch kernels (byte offsets from the 128-byte aligned base)
struct PatchSmem {
    static constexpr int TW = 32, TH = 8, SWD = 8;   // tile, strip width
    static constexpr int FSTRIP = 16 * TH * SWD;      // floats per filter strip
    static constexpr int OFF_FLOW = 16 * TH * TW * 4, OFF_BAR = OFF_FLOW + 2 * TH * TW * 4, OFF_IMG = OFF_BAR + 128;
};

// One SOURCE of a patch tile: stage its flow / filter strips and its image box, compute this
// thread's two pixels into res[k][c] (nothing is stored).  bars[0..2] are initialised mbarriers whose
// next completion has parity `parity`; for a second source the caller passes parity ^ 1.
// `first` = false: the shared memory still holds the previous source, so its readers are drained
// (block barrier) before the TMA engine overwrites it.
// STORE: results go straight to p.outp as they are produced (keeps them out of the register budget)
// and res[][] is left untouched.
template <int C, int SW, int SH, bool STORE, bool FIRST>
__device__ __forceinline__ void patch_pass(unsigned char* sm, const CUtensorMap* m_flow, const CUtensorMap* m_filt,
                                           const CUtensorMap* m_img, const FiArgs& p, float (&res)[2][C]) {
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8, b = blockIdx.z;  // one tile per CTA
    constexpr unsigned parity = FIRST ? 0u : 1u;
    constexpr bool first = FIRST;
    using K = PatchCfg<SW, SH>;
    using S = PatchSmem;
    constexpr int TW = S::TW, TH = S::TH, SWD = S::SWD, FSTRIP = S::FSTRIP;
    float* s_filt = reinterpret_cast<float*>(sm);                            // [4][16][TH][8]
    const float* s_flow = reinterpret_cast<const float*>(sm + S::OFF_FLOW);  // [4][2][TH][8]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::OFF_BAR);           // 0 flow, 1 filter, 2 image
    int* s_bb = reinterpret_cast<int*>(bars + 3);
    const float* s_img = reinterpret_cast<const float*>(sm + S::OFF_IMG);    // [C][SH][SW]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = p.W, H = p.H;

    if (!first) __syncthreads();  // every reader of the previous source's tiles is done
    if (tid == 0) {
        s_bb[0] = INT_MAX; s_bb[1] = INT_MIN; s_bb[2] = INT_MAX; s_bb[3] = INT_MIN;
        if (!first) tma::fence_proxy_async();
        tma::mbar_expect_tx(&bars[0], 2 * TH * TW * 4);
#pragma unroll
        for (int st = 0; st < 4; ++st)
            tma::load_4d(sm + S::OFF_FLOW + st * 2 * TH * SWD * 4, m_flow, x0 + SWD * st, y0, 0, b, &bars[0]);
        tma::mbar_expect_tx(&bars[1], 16 * TH * TW * 4);
#pragma unroll
        for (int st = 0; st < 4; ++st)
            tma::load_4d(sm + st * FSTRIP * 4, m_filt, x0 + SWD * st, y0, 0, b, &bars[1]);
    }
    __syncthreads();
    tma::mbar_wait(&bars[0], parity, 1);

    const int lxx = lane & 7, lyy = lane >> 3;
    const int xl = SWD * warp + lxx;
    const float* flow_s = s_flow + warp * 2 * TH * SWD + lxx;  // + (comp * TH + yl) * 8
    int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int yl = 4 * k + lyy;
        const FiGeom g = fi_geometry(x0 + xl, y0 + yl, W, H, flow_s[yl * SWD], flow_s[(TH + yl) * SWD]);
        if (g.valid && x0 + xl < W && y0 + yl < H) {
            mnx = min(mnx, g.ix); mxx = max(mxx, g.ix);
            mny = min(mny, g.iy); mxy = max(mxy, g.iy);
        }
    }
    mnx = warp_min(mnx); mxx = warp_max(mxx); mny = warp_min(mny); mxy = warp_max(mxy);
    if (lane == 0 && mnx <= mxx) {
        atomicMin(&s_bb[0], mnx); atomicMax(&s_bb[1], mxx);
        atomicMin(&s_bb[2], mny); atomicMax(&s_bb[3], mxy);
    }
    __syncthreads();
    const bool any_valid = s_bb[0] <= s_bb[1];
    int bx = 0, by = 0;
    if (any_valid) {
        bx = s_bb[0] - 1;
        by = s_bb[2] - 1;
        const int need_w = s_bb[1] - s_bb[0] + 4 + 3, need_h = s_bb[3] - s_bb[2] + 4;
        if (need_w > SW) bx += (need_w - SW) / 2;
        if (need_h > SH) by += (need_h - SH) / 2;
        bx = max(0, min(bx, W - SW)) & ~3;
        by = max(0, min(by, H - SH));
    }
    // (the image barrier completes once per pass even when nothing is valid: parities stay in step)
    if (tid == 0) {
        tma::mbar_expect_tx(&bars[2], C * SH * SW * 4);
        tma::load_4d(sm + S::OFF_IMG, m_img, bx, by, 0, b, &bars[2]);
    }
    tma::mbar_wait(&bars[1], parity, 2);
    tma::mbar_wait(&bars[2], parity, 3);

    const float* in1b = p.in1p + b * p.in1.b;
    const float* filt_w = s_filt + warp * FSTRIP;  // this warp's filter strip
    unsigned slow = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int yl = 4 * k + lyy;
        const int x = x0 + xl, y = y0 + yl;
        if (!STORE) {
#pragma unroll
            for (int c = 0; c < C; ++c) res[k][c] = 0.f;
        }
        if (x >= W || y >= H) continue;
        float* outp = p.outp + b * p.out.b + (int64_t)y * p.out.h + x;
        const FiGeom g = fi_geometry(x, y, W, H, flow_s[yl * SWD], flow_s[(TH + yl) * SWD]);
        if (!g.valid) {  // my_lib_kernel.cu:1209-1213: copy the input pixel
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float v = __ldg(in1b + c * p.in1.c + (int64_t)y * p.in1.h + x);
                if (STORE) stg_stream(outp + c * p.out.c, v);
                else res[k][c] = v;
            }
            continue;
        }
        const int lx = g.ix - 1 - bx, ly = g.iy - 1 - by;
        if (!(((unsigned)lx <= (unsigned)(SW - 4)) && ((unsigned)ly <= (unsigned)(SH - 4)))) {
            slow |= 1u << k;
            continue;
        }
        const float a = g.alpha, bt = g.beta;
        const float wTL = (1.0f - a) * (1.0f - bt), wTR = a * (1.0f - bt);
        const float wBL = (1.0f - a) * bt, wBR = a * bt;
        const float* wcol = filt_w + yl * SWD + lxx;  // plane stride TH*8
        float wg[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) wg[t] = wcol[t * TH * SWD];
        const float* base = s_img + ly * SW + lx;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    q[(j >> 1) * 2 + (i >> 1)] =
                        fmaf(base[c * SH * SW + j * SW + i], wg[j * 4 + i], q[(j >> 1) * 2 + (i >> 1)]);
            const float v = wTL * q[0] + wTR * q[1] + wBL * q[2] + wBR * q[3];
            if (STORE) stg_stream(outp + c * p.out.c, v);
            else res[k][c] = v;
        }
    }
    if (__builtin_expect(slow != 0, 0)) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (!((slow >> k) & 1u)) continue;
            const int yl = 4 * k + lyy;
            const int x = x0 + xl, y = y0 + yl;
            const FiGeom g = fi_geometry(x, y, W, H, flow_s[yl * SWD], flow_s[(TH + yl) * SWD]);
            if (STORE) {
                float r[C];
                fwd_slow_pixel<C, K>(p, filt_w + yl * SWD + lxx, TH * SWD, s_img, in1b, g.ix - 1, g.iy - 1, bx, by, g.alpha,
                                     g.beta, r);
                float* outp = p.outp + b * p.out.b + (int64_t)y * p.out.h + x;
#pragma unroll
                for (int c = 0; c < C; ++c) stg_stream(outp + c * p.out.c, r[c]);
            } else {
                fwd_slow_pixel<C, K>(p, filt_w + yl * SWD + lxx, TH * SWD, s_img, in1b, g.ix - 1, g.iy - 1, bx, by, g.alpha,
                                     g.beta, res[k]);
            }
        }
    }
}

__device__ __forceinline__ unsigned char* patch_smem_init() {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((128u - (tma::smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + PatchSmem::OFF_BAR);
    if (threadIdx.x == 0) {
        tma::mbar_init(&bars[0], 1);
        tma::mbar_init(&bars[1], 1);
        tma::mbar_init(&bars[2], 1);
        tma::fence_barrier_init();
    }
    return sm;  // the first patch_pass's block barrier publishes the barriers
}

template <int C, int SW, int SH, int MINB, bool STAGE_OUT>
__global__ void __launch_bounds__(128, MINB)
fi_fwd_patch_kernel(const __grid_constant__ CUtensorMap m_flow, const __grid_constant__ CUtensorMap m_filt,
                    const __grid_constant__ CUtensorMap m_img, const __grid_constant__ CUtensorMap m_out,
                    const __grid_constant__ FiArgs p) {
    constexpr int TW = 32, TH = 8, SWD = 8;
    unsigned char* sm = patch_smem_init();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
    float res[2][C];
    patch_pass<C, SW, SH, !STAGE_OUT, true>(sm, &m_flow, &m_filt, &m_img, p, res);
    const int lxx = lane & 7, lyy = lane >> 3;
    if (STAGE_OUT) {
        // every lane of the warp is done with the warp's filter strip: stage [C][TH][8] there, one
        // TMA store per warp (clipped to the image by the TMA) instead of 4-row partial stores
        float* filt_w = reinterpret_cast<float*>(sm) + warp * PatchSmem::FSTRIP;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c) filt_w[(c * TH + 4 * k + lyy) * SWD + lxx] = res[k][c];
        tma::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma::store_4d(&m_out, x0 + SWD * warp, y0, 0, b, filt_w);
            tma::bulk_commit();
            tma::bulk_wait_read_all();
        }
    }
}

// ------------------------------------------------------------------------------------
// Fused call site (networks/MEMC_Net.py:258-264, MEMC_Net_star.py:272-278):
//     out = occlusion0 * FilterInterpolation(ref0, flow0, filter0)
//         + occlusion1 * FilterInterpolation(ref1, flow1, filter1)
// Two patch passes over the same tile, blended in registers: the two warped frames never go to
// HBM (the composition writes them, reads them back twice and runs three elementwise kernels).
// The blend is rounded exactly like the composition (two products, one sum -- no FMA), so the
// result is bit-identical to it.
// ------------------------------------------------------------------------------------
struct BlendArgs {
    const float* occ0p;
    const float* occ1p;
    View occ0, occ1;  // [B,1,H,W]
};

template <int C, int SW, int SH, int MINB>
__global__ void __launch_bounds__(128, MINB)
fi_blend_patch_kernel(const __grid_constant__ CUtensorMap m_flow0, const __grid_constant__ CUtensorMap m_filt0,
                      const __grid_constant__ CUtensorMap m_img0, const __grid_constant__ CUtensorMap m_flow1,
                      const __grid_constant__ CUtensorMap m_filt1, const __grid_constant__ CUtensorMap m_img1,
                      const __grid_constant__ FiArgs p0, const __grid_constant__ FiArgs p1,
                      const __grid_constant__ BlendArgs bl) {
    constexpr int TW = 32, TH = 8, SWD = 8;
    unsigned char* sm = patch_smem_init();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
    const int lxx = lane & 7, lyy = lane >> 3;
    // the occlusion maps of this thread's pixels: in flight during both passes
    float o0[2], o1[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int x = x0 + SWD * warp + lxx, y = y0 + 4 * k + lyy;
        o0[k] = o1[k] = 0.5f;  // no occlusion maps: the plain mean of MEMC_Net_s (networks/MEMC_Net_s.py:260-264)
        if (bl.occ0p && x < p0.W && y < p0.H) {
            o0[k] = ldg_stream(bl.occ0p + b * bl.occ0.b + (int64_t)y * bl.occ0.h + x);
            o1[k] = ldg_stream(bl.occ1p + b * bl.occ1.b + (int64_t)y * bl.occ1.h + x);
        }
    }
    float r0[2][C], r1[2][C];
    patch_pass<C, SW, SH, false, true>(sm, &m_flow0, &m_filt0, &m_img0, p0, r0);
    patch_pass<C, SW, SH, false, false>(sm, &m_flow1, &m_filt1, &m_img1, p1, r1);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int x = x0 + SWD * warp + lxx, y = y0 + 4 * k + lyy;
        if (x >= p0.W || y >= p0.H) continue;
        float* outp = p0.outp + b * p0.out.b + (int64_t)y * p0.out.h + x;
#pragma unroll
        for (int c = 0; c < C; ++c)
            stg_stream(outp + c * p0.out.c, __fadd_rn(__fmul_rn(o0[k], r0[k][c]), __fmul_rn(o1[k], r1[k][c])));
    }
}

template <int C, int SW, int SH, int MINB>
int launch_blend_patch(cudaStream_t stream, const FiArgs& a0, const FiArgs& a1, const BlendArgs& bl) {
    if (a0.W < SW || a0.H < SH) return 0;
    CUtensorMap m0[5], m1[5];
    if (!make_maps(a0, false, 8, 8, SW, SH, m0) || !make_maps(a1, false, 8, 8, SW, SH, m1)) return 0;
    constexpr size_t smem = (size_t)(16 + 2) * 8 * 32 * 4 + 128 + (size_t)C * SH * SW * 4 + 128;
    if (!ensure_dynamic_smem(fi_blend_patch_kernel<C, SW, SH, MINB>, smem)) return 0;
    dim3 grid((a0.W + 31) / 32, (a0.H + 7) / 8, a0.B);
    fi_blend_patch_kernel<C, SW, SH, MINB><<<grid, 128, smem, stream>>>(m0[0], m0[1], m0[2], m1[0], m1[1], m1[2], a0, a1, bl);
    count_launch();
    return check_launch("FilterInterpolation pair + occlusion blend (fused)") == 0 ? 1 : -1;
}

template <int C, int SW, int SH, int MINB, bool STAGE_OUT>
int launch_fwd_patch(cudaStream_t stream, const FiArgs& a) {
    if (a.W < SW || a.H < SH) return 0;
    CUtensorMap m[5], mout;
    if (!make_maps(a, false, 8, 8, SW, SH, m)) return 0;  // flow / filter boxes are 8-wide strips
    if (!tma::make_map_nchw(&mout, a.outp, a.B, a.C, a.H, a.W, a.out.b, a.out.c, a.out.h, 8, 8, a.C,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE))
        return 0;
    constexpr size_t smem = (size_t)(16 + 2) * 8 * 32 * 4 + 128 + (size_t)C * SH * SW * 4 + 128;
    if (!ensure_dynamic_smem(fi_fwd_patch_kernel<C, SW, SH, MINB, STAGE_OUT>, smem)) return 0;
    dim3 grid((a.W + 31) / 32, (a.H + 7) / 8, a.B);
    fi_fwd_patch_kernel<C, SW, SH, MINB, STAGE_OUT><<<grid, 128, smem, stream>>>(m[0], m[1], m[2], mout, a);
    count_launch();
    return check_launch("FilterInterpolation forward (TMA, 8x4 patches)") == 0 ? 1 : -1;
}

// ====================================================================================
// backward
// ====================================================================================
// TMA stages flow, gradoutput, the 16 filter planes and the image box of a tile; gradinput3 /
// gradinput2 leave through coalesced streaming stores.  gradinput1 -- the scatter -- is
// accumulated in a SHARED-MEMORY box congruent with the image box and flushed once per tile with
// a TMA reduce-add (~1 L2 reduction sector per pixel instead of ~26 with per-tap global atomics).
//
// Shared memory has no native fp32 atomic add on sm_100a (atomicAdd(float*) on shared compiles
// to an ATOMS.CAST.SPIN compare-and-swap loop, measured ~25 cycles per tap); it does have a
// native fire-and-forget int32 add (ATOMS.ADD).  The box is therefore accumulated in FIXED
// POINT with a per-tile power-of-two scale:
//     M     = max over the tile's valid pixels of  max_c|gradoutput_c| * max_t|filter_t|
//             (every contribution is gq * w with |gq| <= |gradoutput_c|, so |contribution| <= M)
//     Kb    = TW*TH: a box cell receives at most one (unclamped) tap per pixel of the tile;
//             border-clamped taps, which could pile up on one cell, bypass the box (global red)
//     scale = 2^e, the largest power of two with  M * Kb * 2^e < 2^31   (no overflow possible)
// Each contribution is rounded to a multiple of 2^-e <= M * Kb * 2^-30: for a 32x8 tile that is
// M * 2^-22, about the fp32 ulp of M, i.e. the absolute precision of adding into an fp32
// accumulator of magnitude M -- and unlike float atomics the tile's sum is order-independent.
// If M is not finite (NaN / Inf gradients) the tile falls back to the float CAS path so that
// NaN/Inf propagate exactly as in the reference.
template <class K>
__host__ __device__ constexpr Layout make_bwd_layout(int C) {
    Layout l{};
    l.off_a = 16 * K::TH * K::TW * 4;                   // gradoutput tile (filter tile sits at 0)
    l.off_flow = l.off_a + C * K::TH * K::TW * 4;
    l.off_bar = l.off_flow + 2 * K::TH * K::TW * 4;
    l.off_img = (l.off_bar + 64 + 127) & ~127;
    l.off_acc = l.off_img + C * K::SH * K::SW * 4;
    l.total = l.off_acc + C * K::SH * K::SW * 4;
    return l;
}

// A pixel whose window touches the image border or leaves the staged box (a few lanes of ~3 % of
// the warps on the benchmark field): per-tap clamping, box or global source / destination per tap.
// Run after the thread's fast pixels (no overlap of live ranges with the fast path); one window row
// per iteration like the fast path.
template <int C, bool OVERWRITE, bool INT_ACC, class K>
__device__ __forceinline__ void bwd_slow_pixel(const FiArgs& p, const float* wcol, const float* gout_px,
                                            const float* s_img, float* s_acc, float scale, const float* in1b,
                                            float* g1b, float* g2, float* g3, int L, int T, int bx, int by, float a,
                                            float bt) {
    constexpr int TW = K::TW, TH = K::TH, SW = K::SW, SH = K::SH;
    const int W = p.W, H = p.H;
    int* s_acci = reinterpret_cast<int*>(s_acc);
    const float gam_y = 1.0f - bt, gam_x = 1.0f - a;
    float gov[C], gq[C][4], q[C][4];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        gov[c] = gout_px[c * TH * TW];
        gq[c][0] = gov[c] * (1.0f - a) * (1.0f - bt);
        gq[c][1] = gov[c] * a * (1.0f - bt);
        gq[c][2] = gov[c] * (1.0f - a) * bt;
        gq[c][3] = gov[c] * a * bt;
        q[c][0] = q[c][1] = q[c][2] = q[c][3] = 0.f;
    }
    // one window row per iteration; its 4 x C source values (box or L2) are loaded before they are used
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
        const bool top = j < 2;
        const int cy = clampi(T + j, 0, H - 1);
        const int uy = cy - by;
        const bool row_in = (unsigned)uy < (unsigned)SH;
        float v[4][C], w[4];
        int so[4];
        bool to_box[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int cx = clampi(L + i, 0, W - 1);
            const int ux = cx - bx;
            const bool in_box = row_in && (unsigned)ux < (unsigned)SW;
            // clamped taps pile up on border cells (up to 9 of one pixel on a corner): they go
            // straight to global memory so that a box cell receives at most ONE tap per pixel
            to_box[i] = in_box && cx == L + i && cy == T + j;
            w[i] = wcol[(j * 4 + i) * TH * TW];
            so[i] = uy * SW + ux;
            const float* gsrc = in1b + (int64_t)cy * p.in1.h + cx;
#pragma unroll
            for (int c = 0; c < C; ++c) v[i][c] = in_box ? s_img[c * SH * SW + so[i]] : __ldg(gsrc + c * p.in1.c);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int h = i >> 1;
            const int cx = clampi(L + i, 0, W - 1);
            float* g1 = g1b + (int64_t)cy * p.gi1.h + cx;
            float a3 = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float gsel = top ? gq[c][h] : gq[c][2 + h];
                if (to_box[i]) {
                    if (INT_ACC) atomicAdd(&s_acci[c * SH * SW + so[i]], __float2int_rn(gsel * w[i] * scale));
                    else atomicAdd(&s_acc[c * SH * SW + so[i]], gsel * w[i]);
                } else {
                    red_add(g1 + c * p.gi1.c, gsel * w[i]);
                }
                a3 = fmaf(gsel, v[i][c], a3);
                const float nq = fmaf(v[i][c], w[i], top ? q[c][h] : q[c][2 + h]);
                q[c][h] = top ? nq : q[c][h];
                q[c][2 + h] = top ? q[c][2 + h] : nq;
            }
            float* dst = g3 + (int64_t)(j * 4 + i) * p.gi3.c;
            if (OVERWRITE) stg_stream(dst, a3);
            else *dst += a3;
        }
    }
    float dx = 0.f, dy = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        dx = fmaf(gov[c], gam_y * (q[c][1] - q[c][0]) + (1.0f - gam_y) * (q[c][3] - q[c][2]), dx);
        dy = fmaf(gov[c], gam_x * (q[c][2] - q[c][0]) + (1.0f - gam_x) * (q[c][3] - q[c][1]), dy);
    }
    stg_stream(g2, dx);
    stg_stream(g2 + p.gi2.c, dy);
}

template <int C, bool OVERWRITE, bool INT_ACC, class K>
__device__ __forceinline__ void bwd_compute_tile(const FiArgs& p, const float* s_filt, const float* s_gout,
                                                 const float* s_flow, const float* s_img, float* s_acc, float scale,
                                                 int x0, int y0, int b, int bx, int by, int lane, int warp) {
    constexpr int TW = K::TW, TH = K::TH, SW = K::SW, SH = K::SH;
    const int W = p.W, H = p.H;
    const float* in1b = p.in1p + b * p.in1.b;
    float* g1b = p.gi1p + b * p.gi1.b;
    int* s_acci = reinterpret_cast<int*>(s_acc);
    static_assert(K::PPT <= 32, "slow-pixel mask");
    unsigned slow = 0;  // bit k: pixel k of this thread needs the clamped path
#pragma unroll 1
    for (int k = 0; k < K::PPT; ++k) {
        int xl, yl;
        tile_pixel<K>(k, lane, warp, xl, yl);
        const int x = x0 + xl, y = y0 + yl;
        if (x >= W || y >= H) continue;
        float* g2 = p.gi2p + b * p.gi2.b + (int64_t)y * p.gi2.h + x;
        float* g3 = p.gi3p + b * p.gi3.b + (int64_t)y * p.gi3.h + x;
        const FiGeom g = fi_geometry(x, y, W, H, s_flow[yl * TW + xl], s_flow[(TH + yl) * TW + xl]);
        if (!g.valid) {  // my_lib_kernel.cu:1256: an invalid pixel contributes nothing
            if (OVERWRITE) {
                stg_stream(g2, 0.f);
                stg_stream(g2 + p.gi2.c, 0.f);
#pragma unroll
                for (int t = 0; t < 16; ++t) stg_stream(g3 + t * p.gi3.c, 0.f);
            }
            continue;
        }
        const float a = g.alpha, bt = g.beta;
        const float gam_y = 1.0f - bt, gam_x = 1.0f - a;  // the reference uses (1 - gamma), not beta
        const int L = g.ix - 1, T = g.iy - 1;
        const int lx = L - bx, ly = T - by;
        // the box lies inside the image (0 <= bx <= W-SW, 0 <= by <= H-SH), so a window inside the
        // box is also inside the image: no per-tap clamping needed
        const bool fast = ((unsigned)lx <= (unsigned)(SW - 4)) && ((unsigned)ly <= (unsigned)(SH - 4));
        const float* wcol = s_filt + yl * TW + xl;  // plane stride TH*TW
        if (__builtin_expect(!fast, 0)) {  // border / outside the staged box: deferred
            slow |= 1u << k;
            continue;
        }
        // Taps outer, channels inner: gradinput3[t] is complete after its channel loop and is
        // stored at once (no 16-register accumulator array), the filter tap is read when needed.
        float gov[C], q[C][4];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            gov[c] = s_gout[(c * TH + yl) * TW + xl];
            q[c][0] = q[c][1] = q[c][2] = q[c][3] = 0.f;
        }
        // top rows (j = 0, 1) feed the TL / TR quadrants, bottom rows (j = 2, 3) BL / BR: two static
        // halves, each a two-iteration row loop that is NOT unrolled (keeps ~12 loads in flight
        // instead of 64); only the half's own quadrant weights are live in its loop
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const float wy = half ? bt : (1.0f - bt);
            float gq[C][2];  // gradoutput x bilinear weight of the (left, right) quadrant of this half
#pragma unroll
            for (int c = 0; c < C; ++c) {
                gq[c][0] = gov[c] * (1.0f - a) * wy;
                gq[c][1] = gov[c] * a * wy;
            }
#pragma unroll 1
            for (int jj = 0; jj < 2; ++jj) {
                const int j = 2 * half + jj;
                const int roff = (ly + j) * SW + lx;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int h = i >> 1;
                    const float w = wcol[(j * 4 + i) * TH * TW];
                    float a3 = 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int o = roff + c * SH * SW + i;
                        const float v = s_img[o];
                        const float gsel = gq[c][h];
                        if (INT_ACC) atomicAdd(&s_acci[o], __float2int_rn(gsel * w * scale));
                        else atomicAdd(&s_acc[o], gsel * w);
                        a3 = fmaf(gsel, v, a3);
                        q[c][2 * half + h] = fmaf(v, w, q[c][2 * half + h]);
                    }
                    float* dst = g3 + (int64_t)(j * 4 + i) * p.gi3.c;
                    if (OVERWRITE) stg_stream(dst, a3);
                    else *dst += a3;  // own pixel: no atomic needed
                }
            }
        }
        float dx = 0.f, dy = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            dx = fmaf(gov[c], gam_y * (q[c][1] - q[c][0]) + (1.0f - gam_y) * (q[c][3] - q[c][2]), dx);
            dy = fmaf(gov[c], gam_x * (q[c][2] - q[c][0]) + (1.0f - gam_x) * (q[c][3] - q[c][1]), dy);
        }
        stg_stream(g2, dx);
        stg_stream(g2 + p.gi2.c, dy);
    }
    if (__builtin_expect(slow != 0, 0)) {
#pragma unroll 1
        for (int k = 0; k < K::PPT; ++k) {
            if (!((slow >> k) & 1u)) continue;
            int xl, yl;
            tile_pixel<K>(k, lane, warp, xl, yl);
            const int x = x0 + xl, y = y0 + yl;
            const FiGeom g = fi_geometry(x, y, W, H, s_flow[yl * TW + xl], s_flow[(TH + yl) * TW + xl]);
            bwd_slow_pixel<C, OVERWRITE, INT_ACC, K>(p, s_filt + yl * TW + xl, s_gout + yl * TW + xl, s_img, s_acc, scale,
                                                     in1b, g1b, p.gi2p + b * p.gi2.b + (int64_t)y * p.gi2.h + x,
                                                     p.gi3p + b * p.gi3.b + (int64_t)y * p.gi3.h + x, g.ix - 1, g.iy - 1,
                                                     bx, by, g.alpha, g.beta);
        }
    }
}

template <int C, bool OVERWRITE, class K>
__global__ void __launch_bounds__(K::NT, K::MINB)
fi_bwd_tma_kernel(const __grid_constant__ CUtensorMap m_flow, const __grid_constant__ CUtensorMap m_gout,
                  const __grid_constant__ CUtensorMap m_filt, const __grid_constant__ CUtensorMap m_img,
                  const __grid_constant__ CUtensorMap m_gi1, const __grid_constant__ FiArgs p) {
    constexpr int TW = K::TW, TH = K::TH, SW = K::SW, SH = K::SH, NT = K::NT;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((128u - (tma::smem_u32(smem_raw) & 127u)) & 127u);
    constexpr Layout lay = make_bwd_layout<K>(C);
    const float* s_filt = reinterpret_cast<const float*>(sm);                 // [16][TH][TW]
    const float* s_gout = reinterpret_cast<const float*>(sm + lay.off_a);     // [C][TH][TW]
    const float* s_flow = reinterpret_cast<const float*>(sm + lay.off_flow);  // [2][TH][TW]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + lay.off_bar);           // 0 flow, 1 gout, 2 filter, 3 image
    int* s_bb = reinterpret_cast<int*>(bars + 4);
    unsigned* s_maxbits = reinterpret_cast<unsigned*>(s_bb + 4);
    const float* s_img = reinterpret_cast<const float*>(sm + lay.off_img);
    float* s_acc = reinterpret_cast<float*>(sm + lay.off_acc);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
    const int W = p.W, H = p.H;

    if (tid == 0) {
        for (int k = 0; k < 4; ++k) tma::mbar_init(&bars[k], 1);
        s_bb[0] = INT_MAX; s_bb[1] = INT_MIN; s_bb[2] = INT_MAX; s_bb[3] = INT_MIN;
        *s_maxbits = 0u;
        tma::fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        tma::mbar_expect_tx(&bars[0], 2 * TH * TW * 4);
        tma::load_4d(sm + lay.off_flow, &m_flow, x0, y0, 0, b, &bars[0]);
        tma::mbar_expect_tx(&bars[1], C * TH * TW * 4);
        tma::load_4d(sm + lay.off_a, &m_gout, x0, y0, 0, b, &bars[1]);
        tma::mbar_expect_tx(&bars[2], 16 * TH * TW * 4);
        tma::load_4d(sm, &m_filt, x0, y0, 0, b, &bars[2]);
    }
    // zero the accumulation box while the loads fly (int 0 and float 0 share the bit pattern)
    {
        float4* a4 = reinterpret_cast<float4*>(s_acc);
        for (int i = tid; i < C * SH * SW / 4; i += NT) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    tma::mbar_wait(&bars[0], 0, 11);
    int bx, by;
    const bool any_valid = tile_box<K>(s_flow, s_bb, x0, y0, W, H, lane, warp, bx, by);  // syncs: box is zeroed
    if (tid == 0 && any_valid) {
        tma::mbar_expect_tx(&bars[3], C * SH * SW * 4);
        tma::load_4d(sm + lay.off_img, &m_img, bx, by, 0, b, &bars[3]);
    }
    tma::mbar_wait(&bars[1], 0, 12);
    tma::mbar_wait(&bars[2], 0, 13);

    // ---- per-tile fixed-point scale.  |x| of a float orders like its bit pattern, and NaN patterns
    // sort above +Inf, so the maxima are integer maxima of (bits & 0x7fffffff): NaN propagates.
    unsigned mbits = 0u;
#pragma unroll
    for (int k = 0; k < K::PPT; ++k) {
        int xl, yl;
        tile_pixel<K>(k, lane, warp, xl, yl);
        const FiGeom g = fi_geometry(x0 + xl, y0 + yl, W, H, s_flow[yl * TW + xl], s_flow[(TH + yl) * TW + xl]);
        if (g.valid && x0 + xl < W && y0 + yl < H) {
            unsigned mg = 0u, mw = 0u;
#pragma unroll
            for (int c = 0; c < C; ++c) mg = max(mg, __float_as_uint(s_gout[(c * TH + yl) * TW + xl]) & 0x7fffffffu);
#pragma unroll
            for (int t = 0; t < 16; ++t) mw = max(mw, __float_as_uint(s_filt[(t * TH + yl) * TW + xl]) & 0x7fffffffu);
            // Inf * 0 = NaN is fine here: any non-finite tile takes the float path
            mbits = max(mbits, __float_as_uint(__uint_as_float(mg) * __uint_as_float(mw)) & 0x7fffffffu);
        }
    }
    {
        const unsigned mb = __reduce_max_sync(0xffffffffu, mbits);
        if (lane == 0 && mb) atomicMax(s_maxbits, mb);
    }
    __syncthreads();
    const float M = __uint_as_float(*s_maxbits);
    const bool finite = *s_maxbits < 0x7f800000u;
    float scale = 0.f, inv_scale = 0.f;
    if (finite && M > 0.f && !(p.flags & MEMC_B200_FLOAT_ACCUM)) {
        int ex;
        frexpf(M, &ex);  // M < 2^ex
        constexpr int LOG2_PX = 31 - __builtin_clz(TW * TH - 1) + 1;  // ceil(log2(TW*TH))
        int e = 31 - ex - LOG2_PX;
        e = max(-120, min(e, 120));
        scale = ldexpf(1.0f, e);
        inv_scale = ldexpf(1.0f, -e);
    }
    if (any_valid) tma::mbar_wait(&bars[3], 0, 14);

    if (scale > 0.f)
        bwd_compute_tile<C, OVERWRITE, true, K>(p, s_filt, s_gout, s_flow, s_img, s_acc, scale, x0, y0, b, bx, by, lane, warp);
    else
        bwd_compute_tile<C, OVERWRITE, false, K>(p, s_filt, s_gout, s_flow, s_img, s_acc, 1.0f, x0, y0, b, bx, by, lane, warp);

    // ---- flush the accumulation box: one TMA reduce-add per tile (clipped to the image by the TMA)
    __syncthreads();
    if (scale > 0.f) {  // fixed point -> fp32 in place
        int4* ai = reinterpret_cast<int4*>(s_acc);
        float4* af = reinterpret_cast<float4*>(s_acc);
        for (int i = tid; i < C * SH * SW / 4; i += NT) {
            const int4 q = ai[i];
            af[i] = make_float4((float)q.x * inv_scale, (float)q.y * inv_scale, (float)q.z * inv_scale, (float)q.w * inv_scale);
        }
    }
    tma::fence_proxy_async();  // generic-proxy writes to s_acc -> visible to the async proxy
    __syncthreads();
    if (tid == 0 && any_valid) {
        tma::reduce_add_4d(&m_gi1, bx, by, 0, b, s_acc);
        tma::bulk_commit();
        tma::bulk_wait_read_all();  // shared memory must stay alive until the TMA has read it
    }
}

// ------------------------------------------------------------------------------- launch
bool make_maps(const FiArgs& a, bool bwd, int TW, int TH, int SW, int SH, CUtensorMap* m) {
    if (!tma::make_map_nchw(&m[0], a.flowp, a.B, 2, a.H, a.W, a.flow.b, a.flow.c, a.flow.h, TW, TH, 2,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
        return false;
    if (!bwd) {
        if (!tma::make_map_nchw(&m[1], a.filtp, a.B, 16, a.H, a.W, a.filt.b, a.filt.c, a.filt.h, TW, TH, 16,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
            return false;
    } else {
        if (!tma::make_map_nchw(&m[1], a.goutp, a.B, a.C, a.H, a.W, a.out.b, a.out.c, a.out.h, TW, TH, a.C,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
            return false;
        if (!tma::make_map_nchw(&m[3], a.gi1p, a.B, a.C, a.H, a.W, a.gi1.b, a.gi1.c, a.gi1.h, SW, SH, a.C,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE))
            return false;
        if (!tma::make_map_nchw(&m[4], a.filtp, a.B, 16, a.H, a.W, a.filt.b, a.filt.c, a.filt.h, TW, TH, 16,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
            return false;
    }
    return tma::make_map_nchw(&m[2], a.in1p, a.B, a.C, a.H, a.W, a.in1.b, a.in1.c, a.in1.h, SW, SH, a.C,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
}

template <int C, class K>
int launch_fwd(cudaStream_t stream, const FiArgs& a) {
    if (a.W < K::SW || a.H < K::SH) return 0;
    CUtensorMap m[5];
    if (!make_maps(a, false, K::TW, K::TH, K::SW, K::SH, m)) return 0;
    constexpr size_t smem = (size_t)make_layout<K>(C, 16, false).total + 128;
    if (!ensure_dynamic_smem(fi_fwd_tma_kernel<C, K>, smem)) return 0;
    dim3 grid((a.W + K::TW - 1) / K::TW, (a.H + K::TH - 1) / K::TH, a.B);
    fi_fwd_tma_kernel<C, K><<<grid, K::NT, smem, stream>>>(m[0], m[1], m[2], a);
    count_launch();
    return check_launch("FilterInterpolation forward (TMA)") == 0 ? 1 : -1;
}

template <int C, bool OW, class K>
int launch_bwd(cudaStream_t stream, const FiArgs& a) {
    if (a.W < K::SW || a.H < K::SH) return 0;
    CUtensorMap m[5];
    if (!make_maps(a, true, K::TW, K::TH, K::SW, K::SH, m)) return 0;
    constexpr size_t smem = (size_t)make_bwd_layout<K>(C).total + 128;
    if (!ensure_dynamic_smem(fi_bwd_tma_kernel<C, OW, K>, smem)) return 0;
    dim3 grid((a.W + K::TW - 1) / K::TW, (a.H + K::TH - 1) / K::TH, a.B);
    fi_bwd_tma_kernel<C, OW, K><<<grid, K::NT, smem, stream>>>(m[0], m[1], m[4], m[2], m[3], a);
    count_launch();
    return check_launch("FilterInterpolation backward (TMA)") == 0 ? 1 : -1;
}

// tile configurations (sweeps: profiles/r01_fi_tile_sweep.md)
//                  TW  TH  SW  SH   NT  MINB
using FwdE6 = Cfg<32, 8, 64, 24, 128, 6>;         // row segments, 37 KB, registers capped at 80: 6 CTAs / SM
using FwdK3 = Cfg<32, 16, 72, 32, 256, 2, true>;  // channel-chunked, 8x4 patches, box pitch 72: 110 KB
using FWD_DEFAULT = FwdE6;
using BwdF = Cfg<32, 8, 64, 28, 256, 3>;          // 64 KB: 3 CTAs / SM, 1 px / thread
using BWD_DEFAULT = BwdF;

}  // namespace

int fi_backward_rows(cudaStream_t stream, const FiArgs& a, bool overwrite);  // filter_interpolation_bwd_rows.cu
int fi_forward_cols(cudaStream_t stream, const FiArgs& a);  // filter_interpolation_fwd_cols.cu (C > 4)

// Kernel selection.  `variant` = MEMC_B200_VARIANT field of the call's flags: 0 is production, the others keep
// earlier kernels reachable for A/B measurements (tools/kbench.py) and cross-checks in the tests.
int fi_forward_fast(cudaStream_t stream, const FiArgs& a) {
    const int variant = (a.flags >> 16) & 0xff;
    if (a.fs != 4 || a.C < 1 || a.W % 4 || a.B > 65535) return 0;
    if (a.C > CB) {  // e.g. the 64-channel context warps of MEMC_Net_star: channel-chunked kernels
        if (variant == 1) return 0;   // generic kernel
        if (variant != 2) {           // production: (pixel, tap column) lanes
            const int r = fi_forward_cols(stream, a);
            if (r != 0) return r;
        }
        return launch_fwd_chunked<FwdK3>(stream, a);  // round-1 kernel (8x4 patches); also the fallback
    }
    switch (a.C) {
        case 1: return launch_fwd<1, FWD_DEFAULT>(stream, a);
        case 2: return launch_fwd<2, FWD_DEFAULT>(stream, a);
        case 3:
            if (variant == 1) return launch_fwd<3, FWD_DEFAULT>(stream, a);       // row segments
            if (variant == 2) return launch_fwd_patch<3, 72, 22, 6, true>(stream, a);  // patches + TMA-staged output
            return launch_fwd_patch<3, 72, 22, 6, false>(stream, a);              // 8x4 patches
        case 4: return launch_fwd<4, FWD_DEFAULT>(stream, a);
    }
    return 0;
}

// fused pair + blend: 1 = handled, 0 = preconditions not met (caller composes the plain ops), -1 = error
int fi_blend_forward_fast(cudaStream_t stream, const FiArgs& a0, const FiArgs& a1, const float* occ0, View v_occ0,
                          const float* occ1, View v_occ1) {
    if (a0.fs != 4 || a0.C != 3 || a0.W % 4 || a0.B > 65535) return 0;
    BlendArgs bl{occ0, occ1, v_occ0, v_occ1};
    return launch_blend_patch<3, 72, 22, 5>(stream, a0, a1, bl);
}

int fi_backward_fast(cudaStream_t stream, const FiArgs& a, bool ow) {
    const int variant = (a.flags >> 16) & 0xff;
    if (a.C > CB) return variant == 1 ? 0 : fi_backward_chunked(stream, a, ow);  // channel chunks (variant 1: generic kernel)
    if (a.fs != 4 || a.C < 1 || a.W % 4 || a.B > 65535) return 0;
    if (variant != 1) {  // production: (pixel, tap row) lanes
        const int r = fi_backward_rows(stream, a, ow);
        if (r != 0) return r;
    }
    switch (a.C) {  // round-1 kernel: one pixel per lane
        case 1: return ow ? launch_bwd<1, true, BWD_DEFAULT>(stream, a) : launch_bwd<1, false, BWD_DEFAULT>(stream, a);
        case 2: return ow ? launch_bwd<2, true, BWD_DEFAULT>(stream, a) : launch_bwd<2, false, BWD_DEFAULT>(stream, a);
        case 3: return ow ? launch_bwd<3, true, BWD_DEFAULT>(stream, a) : launch_bwd<3, false, BWD_DEFAULT>(stream, a);
        case 4: return ow ? launch_bwd<4, true, BWD_DEFAULT>(stream, a) : launch_bwd<4, false, BWD_DEFAULT>(stream, a);
    }
    return 0;
}

}  // namespace memc
