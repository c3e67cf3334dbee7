// filter_interpolation_fwd_cols.cu -- FilterInterpolation forward for sm_100a, fs = 4, C <= 4:
// the "tap-column lanes" kernel (round 2; reference semantics my_lib_kernel.cu:1087-1218).
//
// The round-1 forward (one pixel per lane, 8x4 pixel patches) sits on the L1 / shared-memory data pipe: 130
// wavefronts per 32 pixels, 46 of them bank-conflict replays of the 48 image-tap loads, because a warp's 32
// windows spread over ~11 x 6 source cells that are not a permutation of the banks
// (profiles/r01_ncu_bench_fi_fwd.txt).  Shared-memory LOADS broadcast: lanes that read the SAME word cost
// nothing extra.  So the lanes of a warp are made to overlap on purpose:
//
//   lane = (pixel of a GX x 2 pixel group, tap COLUMN r): the NL lanes of a pixel own the columns r, r + NL, ...
//   of its 4x4 window and walk the 4 tap rows together.  Neighbouring pixels' windows overlap in 3 of 4 columns,
//   so one warp instruction touches ~14-20 distinct words instead of 32, and with a row pitch of 72 words
//   (= 8 mod 32) rows r, r + 1 of the box are 8 banks apart.  Modelled on the benchmark field
//   (tools/bank_model.py --roles): 18.4 (NL = 4, 4x2 groups) / 25.0 (NL = 2, 8x2 groups) wavefronts per 32 pixels
//   and channel against 31.4 for 8x4 patches; the partial sums of a pixel's lanes meet in 1-2 butterfly shuffles.
//
// The NL lanes of a pixel read NL different filter planes of the same pixel at once: the filter tile comes through
// a rank-5 tensor map that splits the plane index into (tap row j, tap column i) and the image row into
// (y >> 1, y & 1) (tma::make_map_taps_cols): a strip of GX pixels lands as [y>>1][j][i][y&1][x] and a warp's filter
// read is 32 consecutive words.  Geometry is evaluated once per pixel and handed to the pixel's lanes with shuffles;
// every tap of a lane is then a compile-time offset from one shared-memory base address.
#include "filter_interpolation.cuh"
#include "tma_utils.cuh"

namespace memc {

namespace {

constexpr int TW = 32, TH = 8, NT = 128;  // 4 warps; warp w owns tile rows 2w, 2w + 1
constexpr int SW = 72, SH = 22;           // image box: pitch 72 words (= 8 mod 32), 22 rows, one TMA load

template <int C, int NL_>
struct Lay {
    static constexpr int NL = NL_;            // lanes per pixel
    static constexpr int GX = 16 / NL;        // group width (pixels); groups are GX x 2
    static constexpr int PXS = 2 * GX;        // pixels per warp step
    static constexpr int NSTEP = 64 / PXS;    // steps per warp (= strips per tile)
    static constexpr int NK = 4 / NL;         // tap columns per lane: r, r + NL, ...
    static constexpr int STRIP = TH * 16 * GX;  // floats per filter strip [y>>1][j][i][y&1][x]
    static constexpr int JBLK = 8 * GX;       // floats per (y>>1, j) block: [i][y&1][x]
    static constexpr int CH = SH * SW;        // channel stride inside the box (words)
    static constexpr int OFF_FLOW = 16 * TH * TW * 4;
    static constexpr int OFF_BAR = OFF_FLOW + 2 * TH * TW * 4;
    static constexpr int OFF_IMG = OFF_BAR + 128;
    static constexpr int TOTAL = OFF_IMG + C * CH * 4;
    static_assert(NL == 2 || NL == 4, "2 or 4 lanes per pixel");
};

// per-pixel geometry in the owner lane: code >= 0 fast ((ly << 8) | lx), -1 per-tap path, -2 invalid flow (copy the
// input pixel, my_lib_kernel.cu:1209-1213), -3 outside the image
struct PxGeo {
    int code, ix, iy;
    float alpha, beta;
};

template <int C, int NL>
__global__ void __launch_bounds__(NT, 6)
fi_fwd_cols_kernel(const __grid_constant__ CUtensorMap m_flow, const __grid_constant__ CUtensorMap m_filt,
                   const __grid_constant__ CUtensorMap m_img, const __grid_constant__ FiArgs p) {
    using Y = Lay<C, NL>;
    constexpr int GX = Y::GX, PXS = Y::PXS, NSTEP = Y::NSTEP, NK = Y::NK;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((128u - (tma::smem_u32(smem_raw) & 127u)) & 127u);
    const float* s_filt = reinterpret_cast<const float*>(sm);               // NSTEP strips
    const float* s_flow = reinterpret_cast<const float*>(sm + Y::OFF_FLOW);  // [2][TH][TW]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Y::OFF_BAR);           // 0 flow, 1 filter, 2 image
    int* s_bb = reinterpret_cast<int*>(bars + 3);
    const float* s_img = reinterpret_cast<const float*>(sm + Y::OFF_IMG);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
    const int W = p.W, H = p.H;
    const int px = lane % GX, py = (lane / GX) & 1, r = lane / PXS;  // this lane's role inside a step

    if (tid == 0) {
        for (int k = 0; k < 3; ++k) tma::mbar_init(&bars[k], 1);
        s_bb[0] = INT_MAX; s_bb[1] = INT_MIN; s_bb[2] = INT_MAX; s_bb[3] = INT_MIN;
        tma::fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        tma::mbar_expect_tx(&bars[0], 2 * TH * TW * 4);
        tma::load_4d(sm + Y::OFF_FLOW, &m_flow, x0, y0, 0, b, &bars[0]);
        tma::mbar_expect_tx(&bars[1], 16 * TH * TW * 4);
#pragma unroll
        for (int s = 0; s < NSTEP; ++s)
            tma::load_5d(sm + s * Y::STRIP * 4, &m_filt, x0 + GX * s, 0, 0, 4 * b, y0 >> 1, &bars[1]);
    }

    // ---- geometry once per pixel.  The warp's 64 pixels (2 rows) are spread over the lanes so that the PXS pixels
    // of step s sit in PXS different lanes under the same register index: lane l owns, for k = 0, 1, the pixel
    // (px, py) of step 2 (l / PXS) + k.
    tma::mbar_wait(&bars[0], 0, 31);
    PxGeo me[2];
    bool me_valid[2];
    {
        int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int s = 2 * (lane / PXS) + k;
            const int xl = GX * s + px, yl = 2 * warp + py;
            const FiGeom geo = fi_geometry(x0 + xl, y0 + yl, W, H, s_flow[yl * TW + xl], s_flow[(TH + yl) * TW + xl]);
            const bool inside = x0 + xl < W && y0 + yl < H;
            me_valid[k] = geo.valid && inside;
            me[k].ix = geo.ix; me[k].iy = geo.iy; me[k].alpha = geo.alpha; me[k].beta = geo.beta;
            me[k].code = inside ? -2 : -3;
            if (me_valid[k]) {
                mnx = min(mnx, geo.ix); mxx = max(mxx, geo.ix);
                mny = min(mny, geo.iy); mxy = max(mxy, geo.iy);
            }
        }
        mnx = __reduce_min_sync(0xffffffffu, mnx); mxx = __reduce_max_sync(0xffffffffu, mxx);
        mny = __reduce_min_sync(0xffffffffu, mny); mxy = __reduce_max_sync(0xffffffffu, mxy);
        if (lane == 0 && mnx <= mxx) {
            atomicMin(&s_bb[0], mnx); atomicMax(&s_bb[1], mxx);
            atomicMin(&s_bb[2], mny); atomicMax(&s_bb[3], mxy);
        }
    }
    __syncthreads();
    const bool any_valid = s_bb[0] <= s_bb[1];
    int bx = 0, by = 0;
    if (any_valid) {
        // windows span [min ix - 1, max ix + 2] x [min iy - 1, max iy + 2]; a span larger than the box centres it (what
        // it misses takes the per-tap path); x origin rounded down to 4 pixels (TMA: 16-byte coordinates); the box is
        // kept inside the image, so a window inside the box needs no clamping
        const int need_w = s_bb[1] - s_bb[0] + 4 + 3, need_h = s_bb[3] - s_bb[2] + 4;
        bx = s_bb[0] - 1;
        by = s_bb[2] - 1;
        if (need_w > SW) bx += (need_w - SW) / 2;
        if (need_h > SH) by += (need_h - SH) / 2;
        bx = max(0, min(bx, W - SW)) & ~3;  // W >= SW, H >= SH and W % 4 == 0 are launch preconditions
        by = max(0, min(by, H - SH));
    }
    if (tid == 0 && any_valid) {
        tma::mbar_expect_tx(&bars[2], C * Y::CH * 4);
        tma::load_4d(sm + Y::OFF_IMG, &m_img, bx, by, 0, b, &bars[2]);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (!me_valid[k]) continue;
        const int lx = me[k].ix - 1 - bx, ly = me[k].iy - 1 - by;
        const bool fast = (unsigned)lx <= (unsigned)(SW - 4) && (unsigned)ly <= (unsigned)(SH - 4);
        me[k].code = fast ? ((ly << 8) | lx) : -1;
    }
    tma::mbar_wait(&bars[1], 0, 32);
    if (any_valid) tma::mbar_wait(&bars[2], 0, 33);

    const float* in1b = p.in1p + b * p.in1.b;
    const int y = y0 + 2 * warp + py;
    float* const out_lane = p.outp + b * p.out.b + (int64_t)y * p.out.h + x0 + px;  // + GX s: this lane's pixel of step s
    const float* const f_lane = s_filt + warp * 4 * Y::JBLK + lane;                 // + s STRIP + j JBLK + k 32: tap (j, r + NL k)
    const float* const img_lane = s_img + r;                                        // + box offset of tap (0, 0) + NL k
#pragma unroll
    for (int s = 0; s < NSTEP; ++s) {
        const int src = (s >> 1) * PXS + (lane & (PXS - 1));
        const int code = __shfl_sync(0xffffffffu, me[s & 1].code, src);
        const float a = __shfl_sync(0xffffffffu, me[s & 1].alpha, src), bt = __shfl_sync(0xffffffffu, me[s & 1].beta, src);
        int Lc = 0, T = 0;
        if (__builtin_expect(__any_sync(0xffffffffu, code == -1), 0)) {  // rare: someone needs full coordinates
            Lc = __shfl_sync(0xffffffffu, me[s & 1].ix, src) - 1;
            T = __shfl_sync(0xffffffffu, me[s & 1].iy, src) - 1;
        }
        const float* f = f_lane + s * Y::STRIP;
        float sum[C];
#pragma unroll
        for (int c = 0; c < C; ++c) sum[c] = 0.f;
        if (__builtin_expect(code >= 0, 1)) {
            // every tap of the lane is a compile-time offset from one base: no per-tap integer arithmetic
            const float* base = img_lane + (code >> 8) * SW + (code & 255);
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                float top[C], bot[C];
#pragma unroll
                for (int c = 0; c < C; ++c) top[c] = bot[c] = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float w = f[j * Y::JBLK + k * 32];
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float v = base[c * Y::CH + j * SW + NL * k];
                        if (j < 2) top[c] = fmaf(v, w, top[c]);
                        else bot[c] = fmaf(v, w, bot[c]);
                    }
                }
                // column r + NL k belongs to the left (< 2) or right quadrants: (1 - alpha) / alpha; rows 0, 1: (1 - beta)
                const float cw = (NL == 2 ? k == 0 : r < 2) ? (1.0f - a) : a;
                const float wt_ = cw * (1.0f - bt), wb_ = cw * bt;
#pragma unroll
                for (int c = 0; c < C; ++c) sum[c] = fmaf(wt_, top[c], fmaf(wb_, bot[c], sum[c]));
            }
        } else if (code == -1) {
            // window touches the image border or leaves the staged box: per-tap clamping, box or global source
#pragma unroll 1
            for (int k = 0; k < NK; ++k) {
                const int i = r + NL * k;
                const int cx = clampi(Lc + i, 0, W - 1);
                const int ux = cx - bx;
                float top[C], bot[C];
#pragma unroll
                for (int c = 0; c < C; ++c) top[c] = bot[c] = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float w = f[j * Y::JBLK + k * 32];
                    const int cy = clampi(T + j, 0, H - 1);
                    const int uy = cy - by;
                    const bool in_box = (unsigned)uy < (unsigned)SH && (unsigned)ux < (unsigned)SW;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float v = in_box ? s_img[c * Y::CH + uy * SW + ux] : __ldg(in1b + c * p.in1.c + (int64_t)cy * p.in1.h + cx);
                        if (j < 2) top[c] = fmaf(v, w, top[c]);
                        else bot[c] = fmaf(v, w, bot[c]);
                    }
                }
                const float cw = i < 2 ? (1.0f - a) : a;
                const float wt_ = cw * (1.0f - bt), wb_ = cw * bt;
#pragma unroll
                for (int c = 0; c < C; ++c) sum[c] = fmaf(wt_, top[c], fmaf(wb_, bot[c], sum[c]));
            }
        }
        // the NL lanes of a pixel are PXS lanes apart: butterfly over r
#pragma unroll
        for (int c = 0; c < C; ++c) {
            sum[c] += __shfl_xor_sync(0xffffffffu, sum[c], 16);
            if (NL == 4) sum[c] += __shfl_xor_sync(0xffffffffu, sum[c], 8);
        }
        float* outp = out_lane + GX * s;
        if (__builtin_expect(code >= -1, 1)) {
#pragma unroll
            for (int c = 0; c < C; ++c)
                if ((c % NL) == r) stg_stream(outp + c * p.out.c, sum[c]);  // lane r of the pixel writes channels r, r + NL, ...
        } else if (code == -2) {  // my_lib_kernel.cu:1209-1213: an invalid flow copies the input pixel
            const float* inp = in1b + (int64_t)y * p.in1.h + x0 + GX * s + px;
#pragma unroll
            for (int c = 0; c < C; ++c)
                if ((c % NL) == r) stg_stream(outp + c * p.out.c, __ldg(inp + c * p.in1.c));
        }
    }
}

template <int C, int NL>
int launch_cols(cudaStream_t stream, const FiArgs& a) {
    using Y = Lay<C, NL>;
    CUtensorMap m[3];
    if (!tma::make_map_nchw(&m[0], a.flowp, a.B, 2, a.H, a.W, a.flow.b, a.flow.c, a.flow.h, TW, TH, 2,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B) ||
        !tma::make_map_taps_cols(&m[1], a.filtp, a.B, a.H, a.W, a.filt.b, a.filt.c, a.filt.h, Y::GX, TH,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B) ||
        !tma::make_map_nchw(&m[2], a.in1p, a.B, a.C, a.H, a.W, a.in1.b, a.in1.c, a.in1.h, SW, SH, a.C,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
        return 0;
    constexpr size_t smem = (size_t)Y::TOTAL + 128;
    if (!ensure_dynamic_smem(fi_fwd_cols_kernel<C, NL>, smem)) return 0;
    dim3 grid((a.W + TW - 1) / TW, (a.H + TH - 1) / TH, a.B);
    fi_fwd_cols_kernel<C, NL><<<grid, NT, smem, stream>>>(m[0], m[1], m[2], a);
    count_launch();
    return check_launch("FilterInterpolation forward (TMA, tap-column lanes)") == 0 ? 1 : -1;
}

}  // namespace

// 1 = handled, 0 = layout preconditions not met (caller falls back), -1 = launch error.  nl = lanes per pixel (2 / 4)
int fi_forward_cols(cudaStream_t stream, const FiArgs& a, int nl) {
    if (a.fs != 4 || a.C < 1 || a.C > 4 || a.W % 4 || a.H % 2 || a.B > 65535 || a.W < SW || a.H < SH) return 0;
    if (a.B > 1 && a.filt.b != 16 * a.filt.c) return 0;  // the tap map folds the batch into the plane index
    switch (a.C) {
        case 1: return nl == 4 ? launch_cols<1, 4>(stream, a) : launch_cols<1, 2>(stream, a);
        case 2: return nl == 4 ? launch_cols<2, 4>(stream, a) : launch_cols<2, 2>(stream, a);
        case 3: return nl == 4 ? launch_cols<3, 4>(stream, a) : launch_cols<3, 2>(stream, a);
        case 4: return nl == 4 ? launch_cols<4, 4>(stream, a) : launch_cols<4, 2>(stream, a);
    }
    return 0;
}

}  // namespace memc
