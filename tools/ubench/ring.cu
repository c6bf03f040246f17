// Micro-benchmark: the K1-backward mask/gradient pattern with the mask rows landing in a per-warp
// shared-memory ring through 4-byte cp.async (128 B per warp instruction, any alignment), DEPTH batches
// of 10 rows x 2 tensors ahead, against plain loads into registers one batch ahead.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void cp4(float* d, const float* s) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((unsigned)__cvta_generic_to_shared(d)), "l"(s) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void wait_group() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int DEPTH, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) kring(const float* __restrict__ mr, const float* __restrict__ mi,
                                                     float* __restrict__ gr, float* __restrict__ gi, int n, int tmax) {
    extern __shared__ float ring[];                       // [warp][DEPTH][20][32]
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    float* my = ring + (size_t)wl * DEPTH * 20 * 32 + lane;
    const int warp = blockIdx.x * WARPS + wl, nwarps = gridDim.x * WARPS;
    const int tiles_per = tmax / 32;
    const long long total = (long long)n * tiles_per;
    for (long long tile = warp; tile < total; tile += nwarps) {
        const int u = (int)(tile / tiles_per), t0 = (int)(tile % tiles_per) * 32;
        const long long base = (long long)u * 161 * tmax + t0 + lane;
        // 16 batches of 10 rows (rows >= 161 clamp)
        auto issue = [&](int b) {
            float* dst = my + (b % DEPTH) * 20 * 32;
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                const int f = b * 10 + i < 161 ? b * 10 + i : 160;
                cp4(dst + i * 32, mr + base + (long long)f * tmax);
                cp4(dst + (10 + i) * 32, mi + base + (long long)f * tmax);
            }
            commit();
        };
#pragma unroll
        for (int b = 0; b < DEPTH; ++b) issue(b);
        for (int b = 0; b < 16; ++b) {
            wait_group<DEPTH - 1>();
            __syncwarp();
            const float* src = my + (b % DEPTH) * 20 * 32;
            float a[10], c[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) { a[i] = src[i * 32]; c[i] = src[(10 + i) * 32]; }
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                const int f = b * 10 + i < 161 ? b * 10 + i : 160;
                gr[base + (long long)f * tmax] = a[i] * c[i];
                gi[base + (long long)f * tmax] = a[i];
            }
            __syncwarp();
            if (b + DEPTH < 16) issue(b + DEPTH); else commit();
        }
        wait_group<0>();
    }
}
template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS) kreg(const float* __restrict__ mr, const float* __restrict__ mi,
                                                    float* __restrict__ gr, float* __restrict__ gi, int n, int tmax) {
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int warp = blockIdx.x * WARPS + wl, nwarps = gridDim.x * WARPS;
    const int tiles_per = tmax / 32;
    const long long total = (long long)n * tiles_per;
    for (long long tile = warp; tile < total; tile += nwarps) {
        const int u = (int)(tile / tiles_per), t0 = (int)(tile % tiles_per) * 32;
        const long long base = (long long)u * 161 * tmax + t0 + lane;
        float a[10], c[10], a2[10], c2[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) { a[i] = mr[base + (long long)i * tmax]; c[i] = mi[base + (long long)i * tmax]; }
        for (int b = 0; b < 16; ++b) {
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                const int f = (b + 1) * 10 + i < 161 ? (b + 1) * 10 + i : 160;
                a2[i] = mr[base + (long long)f * tmax]; c2[i] = mi[base + (long long)f * tmax];
            }
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                const int f = b * 10 + i < 161 ? b * 10 + i : 160;
                gr[base + (long long)f * tmax] = a[i] * c[i];
                gi[base + (long long)f * tmax] = a[i];
            }
#pragma unroll
            for (int i = 0; i < 10; ++i) { a[i] = a2[i]; c[i] = c2[i]; }
        }
    }
}
template <class K> void time_it(const char* name, K launch, int n, int tmax) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaEventRecord(e0);
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    const double bytes = (double)n * (tmax / 32) * 32 * 161 * 16.0;
    printf("%-48s %.3f ms  %.0f GB/s  (%s)\n", name, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int n = 256, tmax = 1001;
    const size_t elems = (size_t)n * 161 * 1024 + 1024;
    float *mr, *mi, *gr, *gi;
    cudaMalloc(&mr, elems * 4); cudaMalloc(&mi, elems * 4); cudaMalloc(&gr, elems * 4); cudaMalloc(&gi, elems * 4);
    cudaMemset(mr, 1, elems * 4); cudaMemset(mi, 1, elems * 4);
#define RING(D, W, CT) { const int smem = W * D * 20 * 32 * 4; cudaFuncSetAttribute(kring<D, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    time_it("ring depth " #D ", " #W " warps x " #CT " CTAs/SM", [&] { kring<D, W><<<148 * CT, 32 * W, smem>>>(mr, mi, gr, gi, n, tmax); }, n, tmax); }
    time_it("registers 1 batch ahead, 5 warps x 3 CTAs/SM", [&] { kreg<5><<<148 * 3, 160>>>(mr, mi, gr, gi, n, tmax); }, n, tmax);
    time_it("registers 1 batch ahead, 4 warps x 8 CTAs/SM", [&] { kreg<4><<<148 * 8, 128>>>(mr, mi, gr, gi, n, tmax); }, n, tmax);
    RING(2, 5, 3) RING(3, 5, 3) RING(4, 5, 3) RING(2, 3, 5) RING(6, 5, 3) RING(4, 8, 2)
    return 0;
}
