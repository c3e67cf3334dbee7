// runtime.cu -- library info, launch accounting and the strided zero-fill used by the
// MEMC_B200_OVERWRITE entry points.
#include "memc_common.cuh"
#include "tma_utils.cuh"
#include <atomic>
#include <mutex>
#include <string.h>

namespace memc {

// Process-global state is limited to this file and is thread-safe: ctypes releases the GIL around every call, so
// autograd's per-device worker threads, DataParallel replicas or a multi-threaded host pipeline may be inside the
// library at the same time.
static std::atomic<unsigned long long> g_launches{0};
static std::mutex g_mutex;  // guards the smem-attribute table and the lazy creation of the scratch pools

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// ---- tensor-map cache: 256 direct-mapped entries per process, keyed by (device, everything the encoding depends on)
namespace tma {
bool encode_cached(CUtensorMap* map, const MapKey& key) {
    struct Entry { bool used; int dev; MapKey key; CUtensorMap map; };
    static Entry table[256];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    uint64_t h = 1469598103934665603ull;  // FNV-1a over the key bytes
    const unsigned char* kb = reinterpret_cast<const unsigned char*>(&key);
    for (size_t i = 0; i < sizeof(MapKey); ++i) h = (h ^ kb[i]) * 1099511628211ull;
    Entry& e = table[(h ^ (uint64_t)dev) & 255u];
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        if (e.used && e.dev == dev && memcmp(&e.key, &key, sizeof(MapKey)) == 0) {
            *map = e.map;
            return true;
        }
    }
    if (!encode_direct(map, key)) return false;
    std::lock_guard<std::mutex> lock(g_mutex);
    e.used = true; e.dev = dev; e.key = key; e.map = *map;
    return true;
}
}  // namespace tma

DeviceGuard::DeviceGuard(const void* ptr) {
    if (!ptr) return;  // empty tensors: nothing will be launched
    cudaPointerAttributes attr;
    if (cudaGetDevice(&prev) != cudaSuccess || cudaPointerGetAttributes(&attr, ptr) != cudaSuccess ||
        (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged)) {
        cudaGetLastError();
        ok = false;
        return;
    }
    if (attr.type == cudaMemoryTypeDevice && attr.device != prev) {
        if (cudaSetDevice(attr.device) != cudaSuccess) {
            cudaGetLastError();
            ok = false;
            return;
        }
        switched = true;
    }
}

DeviceGuard::~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
}

__global__ void __launch_bounds__(256) zero_rows_kernel(float* p, View v, int C, int H, int W) {
    // grid: (ceil(W/256), H, B*C)
    const int w = blockIdx.x * 256 + threadIdx.x;
    if (w >= W) return;
    const int bc = blockIdx.z;
    const int b = bc / C, c = bc - b * C;
    p[b * v.b + c * v.c + (int64_t)blockIdx.y * v.h + w] = 0.0f;
}

bool ensure_dynamic_smem_impl(const void* kernel, size_t bytes) {
    struct Entry { const void* fn; int dev; };
    static Entry done[512];
    static int n_done = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    std::lock_guard<std::mutex> lock(g_mutex);
    for (int i = 0; i < n_done; ++i)
        if (done[i].fn == kernel && done[i].dev == dev) return true;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (n_done < 512) done[n_done++] = Entry{kernel, dev};
    return true;
}

int zero_fill(cudaStream_t stream, float* p, View v, int B, int C, int H, int W) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    const bool dense = v.h == W && (C == 1 || v.c == (int64_t)H * W) && (B == 1 || v.b == (int64_t)C * H * W);
    if (dense) {
        cudaError_t err = cudaMemsetAsync(p, 0, sizeof(float) * (size_t)B * C * H * W, stream);
        count_launch();
        if (err != cudaSuccess) {
            fprintf(stderr, "memc_b200: memset failed: %s\n", cudaGetErrorString(err));
            return -1;
        }
        return 0;
    }
    dim3 grid((W + 255) / 256, H, B * C);
    zero_rows_kernel<<<grid, 256, 0, stream>>>(p, v, C, H, W);
    count_launch();
    return check_launch("zero_fill");
}

// Library-owned stream-ordered memory pool for scratch memory (FlowProjection's occupancy masks and
// accumulator ring, the unfused fallback of the blend op).  A private pool with a release threshold keeps the
// blocks cached across calls; the device's default pool would hand them back to the OS at every
// synchronisation (measured: 2 ms per call).  One pool per device, created on first use.  The pool keeps at most
// kScratchKeepBytes cached between calls (a 4K batch can allocate more; the excess goes back to the driver at the
// next synchronisation) and memc_b200_scratch_trim() releases everything.
constexpr unsigned long long kScratchKeepBytes = 1ull << 30;
static cudaMemPool_t g_pools[64] = {};

static cudaMemPool_t scratch_pool() {
    cudaMemPool_t* pools = g_pools;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_mutex);
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
        unsigned long long keep = kScratchKeepBytes;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        pools[dev] = pool;
    }
    return pools[dev];
}

void* scratch_alloc(cudaStream_t stream, size_t bytes) {
    void* p = nullptr;
    cudaMemPool_t pool = scratch_pool();
    if (!pool || cudaMallocFromPoolAsync(&p, bytes, pool, stream) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void scratch_free(cudaStream_t stream, void* p) {
    if (p) cudaFreeAsync(p, stream);
}

}  // namespace memc

extern "C" int memc_b200_abi_version(void) { return 1; }

#define MEMC_STR2(x) #x
#define MEMC_STR(x) MEMC_STR2(x)
extern "C" const char* memc_b200_build_info(void) {
    return "libmemc_b200 sm_100a nvcc " MEMC_STR(__CUDACC_VER_MAJOR__) "." MEMC_STR(__CUDACC_VER_MINOR__) "." MEMC_STR(
        __CUDACC_VER_BUILD__) " built " __DATE__;
}

extern "C" unsigned long long memc_b200_launch_count(void) { return memc::g_launches.load(std::memory_order_relaxed); }

extern "C" int memc_b200_scratch_trim(void) {
    std::lock_guard<std::mutex> lock(memc::g_mutex);
    int rc = 0;
    for (int d = 0; d < 64; ++d)
        if (memc::g_pools[d] && cudaMemPoolTrimTo(memc::g_pools[d], 0) != cudaSuccess) {
            cudaGetLastError();
            rc = -1;
        }
    return rc;
}
