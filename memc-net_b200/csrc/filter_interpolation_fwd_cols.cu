// filter_interpolation_fwd_cols.cu -- FilterInterpolation forward for sm_100a, fs = 4, C > 4 (the 64-channel context
// features MEMC_Net_star warps with the RGB frames' flow and filter, networks/MEMC_Net_star.py:280-285; reference
// semantics my_lib_kernel.cu:1087-1218): "tap-column lanes", channel-chunked.
//
//   lane = (pixel of an 8 x 2 pixel group, tap column r in {0, 1}): the 2 lanes of a pixel own the columns r, r + 2 of
//   its 4x4 window and walk the 4 tap rows together.  Shared-memory LOADS broadcast, and neighbouring pixels' windows
//   overlap in 3 of 4 columns, so a warp instruction touches ~20 distinct words instead of 32: with a row pitch of
//   72 words (= 8 mod 32) the measured conflict factor of the tap loads is 1.59 against 1.96 for one pixel per lane
//   (profiles/r02_fi_lane_roles.md has the model and the measurements of the other lane maps that were built).
//
// The 2 lanes of a pixel read 2 different filter planes of the same pixel at once: the filter tile comes through a
// rank-5 tensor map that splits the plane index into (tap row j, tap column i) and the image row into (y >> 1, y & 1)
// (tma::make_map_taps_cols): a strip of 8 pixels lands as [y>>1][j][i][y&1][x] and a warp's filter read is 32
// consecutive words.  Geometry is evaluated once per pixel and handed to the pixel's lanes with shuffles; the filter
// taps are pre-multiplied with their quadrant's bilinear weight and stay in registers while the image streams through
// a ring of 4-channel boxes; every tap is one LDS at a compile-time offset plus one FMA; the channel sums of a chunk
// are reduced TRANSPOSED over the pixel's lanes (2 shuffles per 4 channels) and each lane stores 2 channels
// (full 32-byte sectors).  The ring needs no block-wide barrier: the last warp to finish a box refills it.
#include "filter_interpolation.cuh"
#include "tma_utils.cuh"

namespace memc {

namespace {

constexpr int TW = 32;  // tile width
constexpr int SW = 72;  // image box pitch: 72 words (= 8 mod 32)

// per-pixel geometry in the owner lane: code >= 0 fast ((ly << 8) | lx), -1 per-tap path, -2 invalid flow (copy the
// input pixel, my_lib_kernel.cu:1209-1213), -3 outside the image
struct PxGeo {
    int code, ix, iy;
    float alpha, beta;
};

// ------------------------------------------------------------------------------------
// C > 4 (the 64-channel context features MEMC_Net_star warps with the same flow / filter,
// networks/MEMC_Net_star.py:280-285): same lanes, the image streams through a two-box ring of 4-channel boxes
// while the lane's filter taps, bilinear coefficients and box offsets stay in registers.  Here the tap loads are
// everything (16 per channel and pixel against 16 + 2 per pixel for filter and flow), so the conflict factor of
// the lane map is the run time: 1.15 (4 lanes per pixel) / 1.56 (2) against 1.96 for the round-1 patches.
// The partial sums of a chunk's 4 channels are reduced TRANSPOSED over the pixel's lanes (3 resp. 2 shuffles per
// chunk instead of 8 / 4): each lane ends up with the total of the channel(s) it then stores.
// ------------------------------------------------------------------------------------
constexpr int KTH = 16, KSH = 32;  // tile rows (8 warps x 2 rows), box rows

// NL lanes per pixel, CBK channels per streamed box, NBUF boxes in the ring
template <int NL_, int CBK_, int NBUF_>
struct LayK {
    static constexpr int NL = NL_, CBK = CBK_, NBUF = NBUF_, GX = 16 / NL, PXS = 2 * GX, NSTEP = 64 / PXS, NK = 4 / NL;
    static constexpr int STRIP = KTH * 16 * GX, JBLK = 8 * GX;
    static constexpr int CH = KSH * SW, BOX = CBK * CH;  // words
    // shared memory: [barriers | ring box 0 | ring box 1 ...]; the flow + filter STAGING area (36 KB) aliases the LAST ring
    // box: it is dead once every lane holds its taps in registers, and the ring only reaches that box with the second chunk
    static constexpr int OFF_BAR = 0;
    static constexpr int OFF_IMG = 128;
    static constexpr int OFF_STAGE = OFF_IMG + (NBUF - 1) * BOX * 4;  // filter strips, then the flow tile
    static constexpr int OFF_FLOW = OFF_STAGE + 16 * KTH * TW * 4;
    static constexpr int TOTAL = OFF_IMG + NBUF * BOX * 4;
    static_assert((16 + 2) * KTH * TW * 4 <= BOX * 4, "the staging area must fit one ring box");
    static_assert(CBK == 2 || CBK == 4, "2 or 4 channels per box");
};

// An optional SECOND image warped with the same flow and filter in the same pass (the RGB frame next to its 64-channel
// context features: networks/MEMC_Net_star.py:272-285 call FilterInterpolation twice per reference with identical
// offset / filter).  Its channels are simply further chunks of the ring: flow tile, filter tile, geometry, bounding box and
// the lane's taps are shared.  C2 == 0: no second source.
struct Src2 {
    const float* in1p;
    float* outp;
    View in1, out;
    int C;
};

template <class Y>
__global__ void __launch_bounds__(256, 3)
fi_fwd_cols_chunked_kernel(const __grid_constant__ CUtensorMap m_flow, const __grid_constant__ CUtensorMap m_filt,
                           const __grid_constant__ CUtensorMap m_img, const __grid_constant__ CUtensorMap m_img2,
                           const __grid_constant__ FiArgs p, const __grid_constant__ Src2 q) {
    constexpr int NL = Y::NL, CBK = Y::CBK, NBUF = Y::NBUF, GX = Y::GX, PXS = Y::PXS, NSTEP = Y::NSTEP, NK = Y::NK;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((128u - (tma::smem_u32(smem_raw) & 127u)) & 127u);
    const float* s_filt = reinterpret_cast<const float*>(sm + Y::OFF_STAGE);  // NSTEP strips (aliases the last ring box)
    const float* s_flow = reinterpret_cast<const float*>(sm + Y::OFF_FLOW);   // [2][KTH][TW]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Y::OFF_BAR);            // 0 flow, 1 filter, 2.. image ring
    int* s_bb = reinterpret_cast<int*>(bars + 2 + NBUF);
    int* s_done = s_bb + 4;  // [NBUF] warps that have finished with a ring buffer
    unsigned char* const sm_img0 = sm + Y::OFF_IMG;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * KTH, b = blockIdx.z;
    const int W = p.W, H = p.H;
    int bx_ = 0, by_ = 0;  // box origin (set after the bounding box is known; the lambda below reads them by reference)
    const int nchunk1 = (p.C + CBK - 1) / CBK;            // chunks [0, nchunk1): first source; the rest: second source
    const int nchunk = nchunk1 + (q.C + CBK - 1) / CBK;
    // ring slot `buf` <- chunk `ch`
    auto issue_chunk = [&](int ch, int buf) {
        tma::mbar_expect_tx(&bars[2 + buf], Y::BOX * 4);
        if (ch < nchunk1) tma::load_4d(sm_img0 + buf * Y::BOX * 4, &m_img, bx_, by_, ch * CBK, b, &bars[2 + buf]);
        else tma::load_4d(sm_img0 + buf * Y::BOX * 4, &m_img2, bx_, by_, (ch - nchunk1) * CBK, b, &bars[2 + buf]);
    };
    const int px = lane % GX, py = (lane / GX) & 1, r = lane / PXS;

    if (tid == 0) {
        for (int k = 0; k < 2 + NBUF; ++k) tma::mbar_init(&bars[k], 1);
        for (int k = 0; k < NBUF; ++k) s_done[k] = 0;
        s_bb[0] = INT_MAX; s_bb[1] = INT_MIN; s_bb[2] = INT_MAX; s_bb[3] = INT_MIN;
        tma::fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        tma::mbar_expect_tx(&bars[0], 2 * KTH * TW * 4);
        tma::load_4d(sm + Y::OFF_FLOW, &m_flow, x0, y0, 0, b, &bars[0]);
        tma::mbar_expect_tx(&bars[1], 16 * KTH * TW * 4);
#pragma unroll
        for (int s = 0; s < NSTEP; ++s)
            tma::load_5d(sm + Y::OFF_STAGE + s * Y::STRIP * 4, &m_filt, x0 + GX * s, 0, 0, 4 * b, y0 >> 1, &bars[1]);
    }
    // ---- geometry once per pixel (lane l owns, for k = 0, 1, the pixel (px, py) of step 2 (l / PXS) + k)
    tma::mbar_wait(&bars[0], 0, 41);
    PxGeo me[2];
    bool me_valid[2];
    {
        int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int s = 2 * (lane / PXS) + k;
            const int xl = GX * s + px, yl = 2 * warp + py;
            const FiGeom geo = fi_geometry(x0 + xl, y0 + yl, W, H, s_flow[yl * TW + xl], s_flow[(KTH + yl) * TW + xl]);
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
    int bx = 0, by = 0;
    if (s_bb[0] <= s_bb[1]) {
        const int need_w = s_bb[1] - s_bb[0] + 4 + 3, need_h = s_bb[3] - s_bb[2] + 4;
        bx = s_bb[0] - 1;
        by = s_bb[2] - 1;
        if (need_w > SW) bx += (need_w - SW) / 2;
        if (need_h > KSH) by += (need_h - KSH) / 2;
        bx = max(0, min(bx, W - SW)) & ~3;  // W >= SW, H >= KSH and W % 4 == 0 are launch preconditions
        by = max(0, min(by, H - KSH));
    }
    bx_ = bx;
    by_ = by;
    if (tid == 0) {
        for (int ch = 0; ch < NBUF - 1 && ch < nchunk; ++ch) issue_chunk(ch, ch);  // (the last box still holds the staging)
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (!me_valid[k]) continue;
        const int lx = me[k].ix - 1 - bx, ly = me[k].iy - 1 - by;
        const bool fast = (unsigned)lx <= (unsigned)(SW - 4) && (unsigned)ly <= (unsigned)(KSH - 4);
        me[k].code = fast ? ((ly << 8) | lx) : -1;
    }
    tma::mbar_wait(&bars[1], 0, 42);

    // ---- per-step state of this lane, kept across the channel chunks: the box offset of its first tap and its
    // filter taps PRE-MULTIPLIED with the bilinear weight of their quadrant, (1-a | a) x (1-b | b):
    //     out = sum_taps v * (w * quadrant weight)        (one FMA per tap and channel, nothing else)
    // st[s] >= 0: fast (box offset); -1: per-tap path; -2: invalid flow (copy the input pixel); -3: outside the image
    int st[NSTEP];
    float wv[NSTEP][NK][4];
    bool any_slow = false;
#pragma unroll
    for (int s = 0; s < NSTEP; ++s) {
        const int src = (s >> 1) * PXS + (lane & (PXS - 1));
        const int code = __shfl_sync(0xffffffffu, me[s & 1].code, src);
        const float a = __shfl_sync(0xffffffffu, me[s & 1].alpha, src), bt = __shfl_sync(0xffffffffu, me[s & 1].beta, src);
        st[s] = code >= 0 ? (code >> 8) * SW + (code & 255) + r : code;
        any_slow |= code == -1;
        const float* f = s_filt + s * Y::STRIP + warp * 4 * Y::JBLK + lane;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const float cw = (NL == 2 ? k == 0 : r < 2) ? (1.0f - a) : a;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // the main loop is branch free: an invalid pixel reads box word 0 with weight 0; a pixel on the per-tap
                // path keeps its taps (its main-loop sum is garbage and is discarded by the per-tap loop)
                const float w = f[j * Y::JBLK + k * 32] * (cw * (j < 2 ? 1.0f - bt : bt));
                wv[s][k][j] = code >= -1 ? w : 0.f;
            }
        }
    }
    any_slow = __any_sync(0xffffffffu, any_slow);
    // geometry and taps live in registers now: the staging area becomes the last ring box
    __syncthreads();
    if (tid == 0 && NBUF - 1 < nchunk) {
        tma::fence_proxy_async();
        issue_chunk(NBUF - 1, NBUF - 1);
    }

    const int y = y0 + 2 * warp + py;
    // channel(s) of a chunk this lane ends up with after the transposed reduction
    const int my_c = CBK == 2 ? (r & 1) : NL == 4 ? (((r & 1) << 1) | (r >> 1)) : 2 * r;
    for (int ch = 0, buf = 0, par = 0; ch < nchunk; ++ch) {
        const float* s_img = reinterpret_cast<const float*>(sm_img0 + buf * Y::BOX * 4);
        tma::mbar_wait(&bars[2 + buf], par, 43);
        // this chunk's source: channels [c0, c0 + CBK) of the first or of the second image
        const bool second = ch >= nchunk1;
        const int c0 = (second ? ch - nchunk1 : ch) * CBK, C = second ? q.C : p.C;
        const float* in1b = second ? q.in1p + b * q.in1.b : p.in1p + b * p.in1.b;
        float* outb = second ? q.outp + b * q.out.b : p.outp + b * p.out.b;
        const int64_t in_c = second ? q.in1.c : p.in1.c, in_h = second ? q.in1.h : p.in1.h;
        const int64_t out_c = second ? q.out.c : p.out.c, out_h = second ? q.out.h : p.out.h;
#pragma unroll
        for (int s = 0; s < NSTEP; ++s) {
            float sum[CBK];
#pragma unroll
            for (int c = 0; c < CBK; ++c) sum[c] = 0.f;
            {
                // branch free: straight-line loads + FMAs of all steps can be interleaved by the compiler
                const float* base = s_img + max(st[s], 0);
#pragma unroll
                for (int k = 0; k < NK; ++k)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int c = 0; c < CBK; ++c) sum[c] = fmaf(base[c * Y::CH + j * SW + NL * k], wv[s][k][j], sum[c]);
            }
            if (__builtin_expect(any_slow, 0)) {  // warp-uniform: someone in this warp has a pixel on the per-tap path
                const int src = (s >> 1) * PXS + (lane & (PXS - 1));
                const int Lc = __shfl_sync(0xffffffffu, me[s & 1].ix, src) - 1, T = __shfl_sync(0xffffffffu, me[s & 1].iy, src) - 1;
                if (st[s] == -1) {
#pragma unroll
                    for (int c = 0; c < CBK; ++c) sum[c] = 0.f;  // discard what the branch-free loop made of box word 0
#pragma unroll  // (k must stay a compile-time index: wv lives in registers)
                    for (int k = 0; k < NK; ++k) {
                        const int cx = clampi(Lc + r + NL * k, 0, W - 1);
                        const int ux = cx - bx;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int cy = clampi(T + j, 0, H - 1);
                            const int uy = cy - by;
                            const bool in_box = (unsigned)uy < (unsigned)KSH && (unsigned)ux < (unsigned)SW;
#pragma unroll
                            for (int c = 0; c < CBK; ++c) {
                                float v = 0.f;
                                if (in_box) v = s_img[c * Y::CH + uy * SW + ux];
                                else if (c0 + c < C) v = __ldg(in1b + (int64_t)(c0 + c) * in_c + (int64_t)cy * in_h + cx);
                                sum[c] = fmaf(v, wv[s][k][j], sum[c]);
                            }
                        }
                    }
                }
            }
            // transposed reduction over the NL lanes of the pixel (lanes PXS apart)
            float mine0, mine1 = 0.f;
            if (CBK == 4) {
                const bool hi = NL == 4 ? (r & 1) : (r != 0);  // keeps channels 2, 3 of the chunk; sends 0, 1
                const float k0 = hi ? sum[2 % CBK] : sum[0], k1 = hi ? sum[3 % CBK] : sum[1];
                const float g0 = __shfl_xor_sync(0xffffffffu, hi ? sum[0] : sum[2 % CBK], NL == 4 ? 8 : 16);
                const float g1 = __shfl_xor_sync(0xffffffffu, hi ? sum[1] : sum[3 % CBK], NL == 4 ? 8 : 16);
                mine0 = k0 + g0;
                mine1 = k1 + g1;
                if (NL == 4) {
                    const bool hi2 = (r & 2) != 0;  // keeps the second of its two channels
                    const float g = __shfl_xor_sync(0xffffffffu, hi2 ? mine0 : mine1, 16);
                    mine0 = (hi2 ? mine1 : mine0) + g;
                }
            } else {  // 2 channels per box: lane keeps channel r & 1
                const bool hi = (r & 1) != 0;
                mine0 = (hi ? sum[1] : sum[0]) + __shfl_xor_sync(0xffffffffu, hi ? sum[0] : sum[1], NL == 4 ? 8 : 16);
                if (NL == 4) mine0 += __shfl_xor_sync(0xffffffffu, mine0, 16);
            }
            if (st[s] == -3) continue;  // outside the image
            const int x = x0 + GX * s + px;
            float* outp = outb + (int64_t)y * out_h + x;
            const float* inp = in1b + (int64_t)y * in_h + x;
#pragma unroll
            for (int qq = 0; qq < (CBK == 4 && NL == 2 ? 2 : 1); ++qq) {
                const int c = c0 + my_c + qq;
                if (c >= C || (CBK == 2 && NL == 4 && r >= 2)) break;  // (2 channels over 4 lanes: lanes r, r ^ 2 hold the same total)
                float v = qq ? mine1 : mine0;
                if (st[s] == -2) v = __ldg(inp + (int64_t)c * in_c);  // my_lib_kernel.cu:1209-1213
                stg_stream(outp + (int64_t)c * out_c, v);
            }
        }
        // no block-wide barrier: the LAST warp to finish with this buffer refills it, nobody waits for anybody
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();  // this warp's reads of the buffer are ordered before the count
            if (atomicAdd(&s_done[buf], 1) == 7) {
                s_done[buf] = 0;
                __threadfence_block();
                if (ch + NBUF < nchunk) {
                    tma::fence_proxy_async();
                    issue_chunk(ch + NBUF, buf);
                }
            }
        }
        if (++buf == NBUF) { buf = 0; par ^= 1; }
    }
}

template <class Y>
int launch_cols_chunked(cudaStream_t stream, const FiArgs& a, const Src2& q) {
    constexpr int CBK = Y::CBK;
    CUtensorMap m[4];
    if (!tma::make_map_nchw(&m[0], a.flowp, a.B, 2, a.H, a.W, a.flow.b, a.flow.c, a.flow.h, TW, KTH, 2,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B) ||
        !tma::make_map_taps_cols(&m[1], a.filtp, a.B, a.H, a.W, a.filt.b, a.filt.c, a.filt.h, Y::GX, KTH,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B) ||
        !tma::make_map_nchw(&m[2], a.in1p, a.B, a.C, a.H, a.W, a.in1.b, a.in1.c, a.in1.h, SW, KSH, CBK,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
        return 0;
    m[3] = m[2];
    if (q.C > 0 && !tma::make_map_nchw(&m[3], q.in1p, a.B, q.C, a.H, a.W, q.in1.b, q.in1.c, q.in1.h, SW, KSH, CBK,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
        return 0;
    constexpr size_t smem = (size_t)Y::TOTAL + 128;
    if (!ensure_dynamic_smem(fi_fwd_cols_chunked_kernel<Y>, smem)) return 0;
    dim3 grid((a.W + TW - 1) / TW, (a.H + KTH - 1) / KTH, a.B);
    fi_fwd_cols_chunked_kernel<Y><<<grid, 256, smem, stream>>>(m[0], m[1], m[2], m[3], a, q);
    count_launch();
    return check_launch("FilterInterpolation forward (TMA, tap-column lanes, channel-chunked)") == 0 ? 1 : -1;
}

}  // namespace

// 1 = handled, 0 = layout preconditions not met (caller falls back), -1 = launch error
static int cols_preconditions(const FiArgs& a) {
    if (a.fs != 4 || a.W % 4 || a.H % 2 || a.B > 65535 || a.W < SW || a.H < KSH) return 0;
    if (a.B > 1 && a.filt.b != 16 * a.filt.c) return 0;  // the tap map folds the batch into the plane index
    return 1;
}

// C > 4 only
int fi_forward_cols(cudaStream_t stream, const FiArgs& a) {
    if (a.C <= 4 || !cols_preconditions(a)) return 0;
    return launch_cols_chunked<LayK<2, 4, 2>>(stream, a, Src2{});  // 2 lanes / pixel, 4 channels / box, 2 boxes: 74 KB, 3 CTAs / SM
}

// two images, one flow / filter: out = FI(in1, flow, filter), out2 = FI(in2, flow, filter); any channel counts
int fi_forward_cols_pair(cudaStream_t stream, const FiArgs& a, const float* in2, View v_in2, float* out2, View v_out2, int C2) {
    if (a.C < 1 || C2 < 1 || !cols_preconditions(a)) return 0;
    return launch_cols_chunked<LayK<2, 4, 2>>(stream, a, Src2{in2, out2, v_in2, v_out2, C2});
}

}  // namespace memc
