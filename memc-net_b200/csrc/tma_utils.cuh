// tma_utils.cuh -- thin inline-PTX wrappers around the sm_100a async-copy machinery used
// by the fast paths: mbarrier, cp.async.bulk.tensor (TMA tile loads, SASS UTMALDG) and
// cp.reduce.async.bulk.tensor (TMA reduce-add stores, SASS UTMAREDG), plus the host-side
// tensor-map encoder obtained through cudaGetDriverEntryPoint (no link-time libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace memc {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
// make freshly initialised barriers visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// order generic-proxy accesses to shared memory before subsequent async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// plain arrival (release): a consumer warp hands a ring slot back to the producer
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait may suspend the thread in hardware for up to this many ns before reporting "not
// yet": waiting warps then stop competing for issue slots with the warps that compute
// (without the hint ~40 % of the executed instructions of the first version were wait spins).
constexpr uint32_t kSuspendHintNs = 20000;
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(kSuspendHintNs)
        : "memory");
    return ok != 0;
}
// Bounded wait: a lost transaction becomes a trap (launch error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
        if (spin > (1u << 20)) {
            if ((threadIdx.x & 31) == 0)
                printf("memc_b200: mbarrier %d timed out in block (%d,%d,%d) thread %d\n", tag, blockIdx.x,
                       blockIdx.y, blockIdx.z, threadIdx.x);
            __trap();
        }
    }
}

// 4-D tile load global -> shared, completion signalled on `bar` (complete_tx::bytes)
__device__ __forceinline__ void load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                        uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
// 4-D tile reduce-add shared -> global (element-wise atomic add performed by the L2), bulk-group completion
__device__ __forceinline__ void reduce_add_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                              const void* smem_src) {
    asm volatile(
        "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group"
        " [%0, {%1, %2, %3, %4}], [%5];" ::"l"(reinterpret_cast<uint64_t>(map)),
        "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(smem_src))
        : "memory");
}
// 4-D tile store shared -> global
__device__ __forceinline__ void store_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3, const void* smem_src) {
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group"
        " [%0, {%1, %2, %3, %4}], [%5];" ::"l"(reinterpret_cast<uint64_t>(map)),
        "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(smem_src))
        : "memory");
}
// 5-D variants (the tap-split filter / gradinput3 maps of the backward, make_map_taps below)
__device__ __forceinline__ void load_5d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                        uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void store_5d(const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                         const void* smem_src) {
    asm volatile(
        "cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group"
        " [%0, {%1, %2, %3, %4, %5}], [%6];" ::"l"(reinterpret_cast<uint64_t>(map)),
        "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(smem_src))
        : "memory");
}
__device__ __forceinline__ void reduce_add_5d(const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                              const void* smem_src) {
    asm volatile(
        "cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group"
        " [%0, {%1, %2, %3, %4, %5}], [%6];" ::"l"(reinterpret_cast<uint64_t>(map)),
        "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(smem_src))
        : "memory");
}
// 4-D tile prefetch global -> L2 only (no shared memory, no completion tracking)
__device__ __forceinline__ void prefetch_l2_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the bulk groups have finished READING shared memory (safe to reuse / exit)
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// Encoded maps are cached (runtime.cu): cuTensorMapEncodeTiled costs 1-2 us and a launch needs 3-7 maps, which is a
// fifth of a small-frame call; a network calls the ops with the same buffers (the caching allocator hands the same
// blocks back) and shapes frame after frame.  The key is everything the encoding depends on.
struct MapKey {
    const void* base;
    uint32_t rank, promo, swizzle;
    uint64_t gdim[5], gstr[4];
    uint32_t box[5];
};
inline bool encode_direct(CUtensorMap* map, const MapKey& key) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    const cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t box[5];
    for (int i = 0; i < 5; ++i) { gdim[i] = key.gdim[i]; box[i] = key.box[i]; }
    for (int i = 0; i < 4; ++i) gstr[i] = key.gstr[i];
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, key.rank, const_cast<void*>(key.base), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)key.swizzle, (CUtensorMapL2promotion)key.promo,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
#ifdef MEMC_TMA_NO_CACHE  // standalone tools (tools/tma_probe.cu) do not link runtime.cu
inline bool encode_cached(CUtensorMap* map, const MapKey& key) { return encode_direct(map, key); }
#else
bool encode_cached(CUtensorMap* map, const MapKey& key);
#endif

inline bool encode_f32(CUtensorMap* map, uint32_t rank, const float* base, const cuuint64_t* gdim, const cuuint64_t* gstr,
                       const cuuint32_t* box, CUtensorMapL2promotion promo, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_NONE) {
    MapKey k{};
    k.base = base; k.rank = rank; k.promo = (uint32_t)promo; k.swizzle = (uint32_t)swz;
    for (uint32_t i = 0; i < rank; ++i) { k.gdim[i] = gdim[i]; k.box[i] = box[i]; }
    for (uint32_t i = 0; i + 1 < rank; ++i) k.gstr[i] = gstr[i];
    return encode_cached(map, k);
}

// fp32 NCHW tensor [B,C,H,W] with element strides (b,c,h,1) -> rank-4 tiled map, box (bw,bh,bc,1).
// Returns false when the layout cannot be described (alignment / size limits).
inline bool make_map_nchw(CUtensorMap* map, const float* base, int B, int C, int H, int W, int64_t sb, int64_t sc,
                          int64_t sh, int bw, int bh, int bc, CUtensorMapL2promotion promo) {
    if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) return false;
    // a size-1 dimension's stride is irrelevant: substitute a legal one
    if (H == 1) sh = W;
    if (C == 1) sc = (int64_t)H * sh;
    if (B == 1) sb = (int64_t)C * sc;
    if (sh % 4 || sc % 4 || sb % 4) return false;  // strides must be multiples of 16 bytes
    if (sh < W || sc <= 0 || sb <= 0) return false;
    if (bw > 256 || bh > 256 || bc > 256 || (bw * 4) % 16) return false;
    const cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t gstr[3] = {(cuuint64_t)sh * 4, (cuuint64_t)sc * 4, (cuuint64_t)sb * 4};
    for (int i = 0; i < 3; ++i)
        if (gstr[i] >= (1ull << 40)) return false;
    const cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc, 1u};
    return encode_f32(map, 4, base, gdim, gstr, box, promo);
}

// A [B,16,H,W] tensor of 4x4 tap planes (plane t = 4 j + i: tap row j, tap column i) as a rank-5 map with the
// plane index split into its two digits and the ROW digit innermost after x:
//     dims (W, j, i, H, B), strides (4 sc, sc, sh, sb), box (bw, 4, 4, bh, 1)
// A box lands in shared memory as [y][i][j][x]: the word of (y, i, j, x) is ((y*4 + i)*4 + j)*bw + x, so with
// bw == 8 the four tap ROWS of 8 neighbouring pixels are 32 consecutive words -- one conflict-free wavefront for
// a warp whose lanes are (pixel, tap row) pairs.  (The strides are not monotonic; the TMA does not care.)
inline bool make_map_taps(CUtensorMap* map, const float* base, int B, int H, int W, int64_t sb, int64_t sc, int64_t sh,
                          int bw, int bh, CUtensorMapL2promotion promo) {
    if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) return false;
    if (H == 1) sh = W;
    if (B == 1) sb = 16 * sc;
    if (sh % 4 || sc % 4 || sb % 4) return false;
    if (sh < W || sc <= 0 || sb <= 0) return false;
    if (bw > 256 || bh > 256 || (bw * 4) % 16) return false;
    const cuuint64_t gdim[5] = {(cuuint64_t)W, 4u, 4u, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t gstr[4] = {(cuuint64_t)sc * 16, (cuuint64_t)sc * 4, (cuuint64_t)sh * 4, (cuuint64_t)sb * 4};
    for (int i = 0; i < 4; ++i)
        if (gstr[i] >= (1ull << 40)) return false;
    const cuuint32_t box[5] = {(cuuint32_t)bw, 4u, 4u, (cuuint32_t)bh, 1u};
    return encode_f32(map, 5, base, gdim, gstr, box, promo);
}

// The same [B,16,H,W] tap tensor for (pixel, tap COLUMN) lanes working on groups of gx x 2 pixels (forward):
//     dims (W, y&1, i, 4 b + j, y>>1), strides (sh, sc, 4 sc, 2 sh), box (gx, 2, 4, 4, bh/2)
// lands as [y>>1][j][i][y&1][x]: for one tap row j the words of (i, y&1, x) are 8 gx consecutive words.
// Folding the batch into the j digit needs dense plane batches (sb == 16 sc) and an even H.
inline bool make_map_taps_cols(CUtensorMap* map, const float* base, int B, int H, int W, int64_t sb, int64_t sc, int64_t sh,
                               int gx, int bh, CUtensorMapL2promotion promo) {
    if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) return false;
    if (H % 2 || bh % 2 || (B > 1 && sb != 16 * sc)) return false;
    if (sh % 4 || sc % 4 || sh < W || sc <= 0) return false;
    if (gx > 256 || bh > 512 || (gx * 4) % 16) return false;
    const cuuint64_t gdim[5] = {(cuuint64_t)W, 2u, 4u, (cuuint64_t)4 * B, (cuuint64_t)(H / 2)};
    const cuuint64_t gstr[4] = {(cuuint64_t)sh * 4, (cuuint64_t)sc * 4, (cuuint64_t)sc * 16, (cuuint64_t)sh * 8};
    for (int i = 0; i < 4; ++i)
        if (gstr[i] >= (1ull << 40)) return false;
    const cuuint32_t box[5] = {(cuuint32_t)gx, 2u, 4u, 4u, (cuuint32_t)(bh / 2)};
    return encode_f32(map, 5, base, gdim, gstr, box, promo);
}

}  // namespace tma
}  // namespace memc
