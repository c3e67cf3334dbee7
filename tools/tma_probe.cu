// tma_probe.cu -- which TMA forms does this driver / GPU accept?  (development probe, run on the B200 box)
//   nvcc -gencode arch=compute_100a,code=sm_100a -I memc-net_b200/csrc -o /tmp/tma_probe tools/tma_probe.cu && for t in 0 1 2 3 4 5 6 7; do /tmp/tma_probe $t; done
// Each test runs in its own process (an illegal instruction kills the context).
#define MEMC_TMA_NO_CACHE
#include "tma_utils.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace memc;

__global__ void k_load5(const __grid_constant__ CUtensorMap m, float* out, int x, int y, int b) {
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 8192);
    if (threadIdx.x == 0) { tma::mbar_init(bar, 1); tma::fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { tma::mbar_expect_tx(bar, 4096); tma::load_5d(sm, &m, x, 0, 0, y, b, bar); }
    tma::mbar_wait(bar, 0, 1);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = reinterpret_cast<float*>(sm)[i];
}
__global__ void k_store5(const __grid_constant__ CUtensorMap m, int x, int y, int b, int reduce) {
    extern __shared__ __align__(1024) unsigned char sm[];
    float* s = reinterpret_cast<float*>(sm);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (float)i;
    tma::fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (reduce) tma::reduce_add_5d(&m, x, 0, 0, y, b, s); else tma::store_5d(&m, x, 0, 0, y, b, s);
        tma::bulk_commit(); tma::bulk_wait_all();
    }
}
__global__ void k_load4(const __grid_constant__ CUtensorMap m, float* out, int x, int y, int b, int n, int dst_off) {
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16384);
    if (threadIdx.x == 0) { tma::mbar_init(bar, 1); tma::fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { tma::mbar_expect_tx(bar, n * 4); tma::load_4d(sm + dst_off, &m, x, y, 0, b, bar); }
    tma::mbar_wait(bar, 0, 2);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float*>(sm + dst_off)[i];
}
__global__ void k_red4(const __grid_constant__ CUtensorMap m, int x, int y, int b, int n) {
    extern __shared__ __align__(1024) unsigned char sm[];
    float* s = reinterpret_cast<float*>(sm);
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = 1.0f;
    tma::fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) { tma::reduce_add_4d(&m, x, y, 0, b, s); tma::bulk_commit(); tma::bulk_wait_all(); }
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("  CUDA error: %s\n", cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char** argv) {
    const int t = argc > 1 ? atoi(argv[1]) : 0;
    const int B = 2, H = 24, W = 96;
    std::vector<float> h((size_t)B * 16 * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *out;
    CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMalloc(&out, 1 << 16));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    std::vector<float> ho(16384);
    CUtensorMap m;
    const int64_t sc = (int64_t)H * W, sb = 16 * sc, sh = W;
    if (t <= 2) {
        printf("test %d: rank-5 tap-split map (non-monotonic strides), %s\n", t, t == 0 ? "load" : t == 1 ? "store" : "reduce-add");
        if (!tma::make_map_taps(&m, d, B, H, W, sb, sc, sh, 8, 8, CU_TENSOR_MAP_L2_PROMOTION_NONE)) { printf("  encode FAILED\n"); return 1; }
        printf("  encode ok\n");
        const int x = 16, y = 8, b = 1;
        if (t == 0) {
            k_load5<<<1, 128, 8192 + 64>>>(m, out, x, y, b);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(ho.data(), out, 4096, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int yy = 0; yy < 8; ++yy) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) for (int xx = 0; xx < 8; ++xx) {
                const float exp = h[(size_t)b * sb + (4 * j + i) * sc + (y + yy) * sh + x + xx];
                if (ho[((yy * 4 + i) * 4 + j) * 8 + xx] != exp) ++bad;
            }
            printf("  layout [y][i][j][x]: %d mismatches\n", bad);
        } else {
            k_store5<<<1, 128, 8192>>>(m, x, y, b, t == 2);
            CK(cudaDeviceSynchronize());
            std::vector<float> h2(h.size());
            CK(cudaMemcpy(h2.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0, touched = 0;
            for (size_t k = 0; k < h.size(); ++k) if (h2[k] != h[k]) ++touched;
            for (int yy = 0; yy < 8; ++yy) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) for (int xx = 0; xx < 8; ++xx) {
                const size_t g = (size_t)b * sb + (4 * j + i) * sc + (y + yy) * sh + x + xx;
                const float v = (float)(((yy * 4 + i) * 4 + j) * 8 + xx);
                if (h2[g] != (t == 2 ? h[g] + v : v)) ++bad;
            }
            printf("  %d mismatches, %d elements changed (expect <= 1024)\n", bad, touched);
        }
        return 0;
    }
    // 4-D [B,3,H,W] tests on the first 3 planes
    const int C = 3;
    if (t == 3 || t == 4 || t == 5 || t == 7) {
        const int bw = t == 7 ? 128 : 72, bh = 4;
        printf("test %d: rank-4 box {%d,%d,%d}: %s\n", t, bw, bh, C,
               t == 3 ? "load at negative x / y" : t == 4 ? "load hanging over right / bottom" : t == 5 ? "reduce-add hanging over the edges" : "box wider than the tensor");
        if (!tma::make_map_nchw(&m, d, B, C, H, W, sb, sc, sh, bw, bh, C, CU_TENSOR_MAP_L2_PROMOTION_NONE)) { printf("  encode FAILED\n"); return 1; }
        const int n = bw * bh * C;
        const int x = t == 3 ? -4 : t == 7 ? -8 : 60, y = t == 3 ? -1 : 22;
        if (t == 5) {
            k_red4<<<1, 128, n * 4>>>(m, -4, -2, 1, n);
            CK(cudaDeviceSynchronize());
            k_red4<<<1, 128, n * 4>>>(m, 60, 22, 1, n);
            CK(cudaDeviceSynchronize());
            std::vector<float> h2(h.size());
            CK(cudaMemcpy(h2.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
            int changed = 0, bad = 0;
            for (size_t k = 0; k < h.size(); ++k) { if (h2[k] != h[k]) { ++changed; if (h2[k] != h[k] + 1.0f) ++bad; } }
            printf("  %d elements changed (expect %d), %d wrong\n", changed, C * (68 * 2 + 36 * 2), bad);
            return 0;
        }
        k_load4<<<1, 128, 16384 + 64>>>(m, out, x, y, 1, n, 0);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(ho.data(), out, n * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int c = 0; c < C; ++c) for (int yy = 0; yy < bh; ++yy) for (int xx = 0; xx < bw; ++xx) {
            const int gx = x + xx, gy = y + yy;
            const float exp = (gx < 0 || gy < 0 || gx >= W || gy >= H) ? 0.f : h[(size_t)1 * sb + c * sc + gy * sh + gx];
            if (ho[(c * bh + yy) * bw + xx] != exp) ++bad;
        }
        printf("  %d mismatches (out-of-bounds elements must read 0)\n", bad);
        return 0;
    }
    if (t == 10 || t == 11) {
        printf("test %d: rank-5 %s hanging over the right / bottom edge\n", t, t == 10 ? "store" : "reduce-add");
        if (!tma::make_map_taps(&m, d, B, H, W, sb, sc, sh, 8, 8, CU_TENSOR_MAP_L2_PROMOTION_NONE)) { printf("  encode FAILED\n"); return 1; }
        k_store5<<<1, 128, 8192>>>(m, 92, 20, 1, t == 11);
        CK(cudaDeviceSynchronize());
        std::vector<float> h2(h.size());
        CK(cudaMemcpy(h2.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
        int changed = 0;
        for (size_t k = 0; k < h.size(); ++k) if (h2[k] != h[k]) ++changed;
        printf("  %d elements changed (expect about %d)\n", changed, 16 * 4 * 4);
        return 0;
    }
    if (t == 8 || t == 9) {
        printf("test %d: rank-4 reduce-add box {72,4,3} %s\n", t, t == 8 ? "at negative x / y only" : "hanging over right / bottom only");
        if (!tma::make_map_nchw(&m, d, B, C, H, W, sb, sc, sh, 72, 4, C, CU_TENSOR_MAP_L2_PROMOTION_NONE)) { printf("  encode FAILED\n"); return 1; }
        if (t == 8) k_red4<<<1, 128, 72 * 4 * C * 4>>>(m, -4, -2, 1, 72 * 4 * C);
        else k_red4<<<1, 128, 72 * 4 * C * 4>>>(m, 60, 22, 1, 72 * 4 * C);
        CK(cudaDeviceSynchronize());
        std::vector<float> h2(h.size());
        CK(cudaMemcpy(h2.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
        int changed = 0, bad = 0;
        for (size_t k = 0; k < h.size(); ++k) { if (h2[k] != h[k]) { ++changed; if (h2[k] != h[k] + 1.0f) ++bad; } }
        printf("  %d elements changed (expect %d), %d wrong\n", changed, t == 8 ? C * 68 * 2 : C * 36 * 2, bad);
        return 0;
    }
    if (t == 6) {
        printf("test 6: rank-4 load to a 16-byte (not 128-byte) aligned shared address\n");
        if (!tma::make_map_nchw(&m, d, B, C, H, W, sb, sc, sh, 32, 4, 1, CU_TENSOR_MAP_L2_PROMOTION_NONE)) { printf("  encode FAILED\n"); return 1; }
        k_load4<<<1, 128, 16384 + 64>>>(m, out, 8, 3, 0, 128, 32);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(ho.data(), out, 512, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int yy = 0; yy < 4; ++yy) for (int xx = 0; xx < 32; ++xx) if (ho[yy * 32 + xx] != h[(3 + yy) * sh + 8 + xx]) ++bad;
        printf("  %d mismatches\n", bad);
        return 0;
    }
    return 0;
}
