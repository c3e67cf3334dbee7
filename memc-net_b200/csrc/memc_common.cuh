// memc_common.cuh -- shared device/host helpers for libmemc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define MEMC_B200_STREAM_T
typedef cudaStream_t memc_stream_t;
#include "../../include/memc_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmemc_b200 is written for sm_100a (Blackwell B200) only"
#endif

namespace memc {

// one NCHW fp32 tensor view: base pointer + element strides (w-stride == 1)
struct View {
    int64_t b, c, h;
};
static inline View mk_view(memc_strides s) { return View{s.b, s.c, s.h}; }
static inline View mk_view(int b, int c, int h) { return View{(int64_t)b, (int64_t)c, (int64_t)h}; }

// dense == planes are back to back and rows are W long: enables memset / flat indexing
static inline bool is_dense(View v, int C, int H, int W) {
    return v.h == W && v.c == (int64_t)H * W && v.b == (int64_t)C * H * W;
}
// every stride and the base pointer allow 16-byte vector access along w
static inline bool vec4_ok(View v, const void* p, int W) {
    return (W % 4 == 0) && (v.h % 4 == 0) && (v.c % 4 == 0) && (v.b % 4 == 0) &&
           ((reinterpret_cast<uintptr_t>(p) & 15u) == 0);
}

void count_launch(int n = 1);  // runtime.cu: atomic counter behind memc_b200_launch_count()

// The launchers run on whatever device OWNS the operands, not on whatever device happens to be current: the
// guard makes that device current for the duration of the call and restores the caller's on exit (a tensor on
// cuda:1 with cuda:0 current would otherwise be launched on GPU 0 against GPU-1 pointers).  `ok` is false when
// the pointer is not device memory (the entry points then return -1, as for any layout they cannot take).
struct DeviceGuard {
    int prev = -1;
    bool switched = false, ok = true;
    explicit DeviceGuard(const void* ptr);
    ~DeviceGuard();
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// reference convention: print and return -1 on a launch error (my_lib_kernel.cu:1558-1564)
static inline int check_launch(const char* what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        fprintf(stderr, "memc_b200: launch failed in %s: %s\n", what, cudaGetErrorString(err));
        return -1;
    }
    return 0;
}

// Opt a kernel in to `bytes` of dynamic shared memory once per device (function attributes are
// per device; one process may drive several).  Returns false if the device refuses.
bool ensure_dynamic_smem_impl(const void* kernel, size_t bytes);  // runtime.cu: table keyed by (kernel, device)
template <class KernelT>
inline bool ensure_dynamic_smem(KernelT* kernel, size_t bytes) {
    return ensure_dynamic_smem_impl(reinterpret_cast<const void*>(kernel), bytes);
}

// stream-ordered scratch memory from the library's private pool (runtime.cu); nullptr on failure
void* scratch_alloc(cudaStream_t stream, size_t bytes);
void scratch_free(cudaStream_t stream, void* p);

// zero-fill a [B,C,H,W] strided tensor on `stream` (memset when dense, kernel otherwise)
int zero_fill(cudaStream_t stream, float* p, View v, int B, int C, int H, int W);

// ------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------
// streaming (read-once) global loads: keep them out of L1 so the gathered image tile stays
__device__ __forceinline__ float ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
// write-once global stores: streaming hint
__device__ __forceinline__ void stg_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void stg_stream4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

// fire-and-forget float add (REDG.E.ADD.F32)
__device__ __forceinline__ void red_add(float* p, float v) { atomicAdd(p, v); }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
// max that PROPAGATES NaN (fmaxf drops it): used to detect non-finite tiles
__device__ __forceinline__ float fmaxf_nan(float a, float b) { return (a != a || b != b) ? a + b : fmaxf(a, b); }

// Geometry of the adaptive warp at one output pixel; decisions in fp32 exactly as the
// reference (my_lib_kernel.cu:1126-1138): truncation, validity incl. |flow| < extent/2.
struct FiGeom {
    bool valid;
    int ix, iy;
    float alpha, beta;
};
__device__ __forceinline__ FiGeom fi_geometry(int w, int h, int W, int H, float fx, float fy) {
    FiGeom g;
    const float x2 = (float)w + fx;
    const float y2 = (float)h + fy;
    g.valid = (x2 >= 0.0f) && (y2 >= 0.0f) && (x2 <= (float)(W - 1)) && (y2 <= (float)(H - 1)) &&
              (fabsf(fx) < (float)W / 2.0f) && (fabsf(fy) < (float)H / 2.0f);
    g.ix = g.valid ? (int)x2 : 0;
    g.iy = g.valid ? (int)y2 : 0;
    g.alpha = x2 - (float)g.ix;
    g.beta = y2 - (float)g.iy;
    return g;
}

}  // namespace memc
