// Micro-benchmark: cost of staging a 33x160-sample tile into frame columns with different copy
// flavours.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldgsts ldgsts.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
__device__ __forceinline__ void cp8(void* d, const void* s) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(d)), "l"(s) : "memory");
}
__device__ __forceinline__ void cp4(void* d, const void* s) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((unsigned)__cvta_generic_to_shared(d)), "l"(s) : "memory");
}
__device__ __forceinline__ void cp16(void* d, const void* s) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(d)), "l"(s) : "memory");
}
__device__ __forceinline__ void waitall() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ int slot_of(int j) { return ((3 * (j % 5)) % 5) * 32 + ((13 * (j & 31)) & 31); }

// MODE 0: cp.async 8B, dst slot-major pitch 33 (scattered over 32 lines)
// MODE 1: cp.async 8B, dst frame-major pitch 161 (scattered inside 1288 B)
// MODE 2: cp.async 8B, dst contiguous (row-major raw tile)
// MODE 3: LDG.64 + STS.64 x2, slot-major (all loads of the share first)
// MODE 4: cp.async 16B contiguous raw tile
// MODE 5: cp.async 4B x2 scattered slot-major (lanes along sample)
template <int MODE>
__global__ void __launch_bounds__(96, 5) k(const float* __restrict__ wave, long long* out, int iters, int tiles) {
    extern __shared__ __align__(16) float2 S[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int sa[3], sb[3];
    for (int q = 0; q < 3; ++q) {
        int c = lane + 32 * q; if (c >= 80) c = 0;
        if (MODE == 1) { sa[q] = slot_of(c); sb[q] = slot_of(c + 80); }
        else           { sa[q] = slot_of(c) * 33; sb[q] = slot_of(c + 80) * 33; }
    }
    long long t_issue = 0, t_total = 0;
    for (int it = 0; it < iters; ++it) {
        const int tile = (blockIdx.x + it * gridDim.x) % tiles;
        const float* base = wave + (long long)tile * 32 * 160;
        __syncthreads();
        const long long c0 = clock64();
        if (MODE == 0 || MODE == 1 || MODE == 5) {
            for (int r = w * 11; r < w * 11 + 11; ++r) {
                const float* src = base + r * 160 + 2 * lane;
                #pragma unroll
                for (int q = 0; q < 3; ++q) {
                    if (q == 2 && lane >= 16) break;
                    float2* da = MODE == 1 ? S + (r < 32 ? r : 0) * 161 + sa[q] : S + sa[q] + (r < 32 ? r : 0);
                    float2* db = MODE == 1 ? S + (r >= 1 ? r - 1 : 0) * 161 + sb[q] : S + sb[q] + (r >= 1 ? r - 1 : 0);
                    if (MODE == 5) { cp4(da, src + 64 * q); cp4((float*)da + 1, src + 64 * q + 1); cp4(db, src + 64 * q); cp4((float*)db + 1, src + 64 * q + 1); }
                    else { cp8(da, src + 64 * q); cp8(db, src + 64 * q); }
                }
            }
        } else if (MODE == 2) {
            for (int i = threadIdx.x; i < 33 * 80; i += 96) cp8(S + i, base + 2 * i);
        } else if (MODE == 4) {
            for (int i = threadIdx.x; i < 33 * 40; i += 96) cp16((float4*)S + i, base + 4 * i);
        } else if (MODE == 3) {
            float2 v[11][3];
            const float2* src = reinterpret_cast<const float2*>(base + w * 11 * 160) + lane;
            #pragma unroll
            for (int i = 0; i < 11; ++i) { v[i][0] = __ldg(src + i * 80); v[i][1] = __ldg(src + i * 80 + 32); if (lane < 16) v[i][2] = __ldg(src + i * 80 + 64); }
            #pragma unroll
            for (int i = 0; i < 11; ++i) {
                const int r = w * 11 + i;
                #pragma unroll
                for (int q = 0; q < 3; ++q) {
                    if (q == 2 && lane >= 16) break;
                    S[sa[q] + (r < 32 ? r : 0)] = v[i][q];
                    S[sb[q] + (r >= 1 ? r - 1 : 0)] = v[i][q];
                }
            }
        }
        const long long c1 = clock64();
        waitall();
        __syncthreads();
        const long long c2 = clock64();
        t_issue += c1 - c0; t_total += c2 - c0;
    }
    if (lane == 0) { out[(blockIdx.x * 3 + w) * 2] = t_issue; out[(blockIdx.x * 3 + w) * 2 + 1] = t_total; }
    if (S[threadIdx.x].x == 1.2345f) out[0] = 0;
}
template <int MODE> void run(const float* wave, long long* out, int tiles, const char* name) {
    const int smem = 160 * 33 * 8 + 3088, grid = 148 * 5, iters = 20;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 96, smem>>>(wave, out, 2, tiles);
    cudaEventRecord(e0);
    k<MODE><<<grid, 96, smem>>>(wave, out, iters, tiles);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid * 3 * 2);
    cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost);
    double si = 0, st = 0; for (int i = 0; i < grid * 3; ++i) { si += h[2 * i]; st += h[2 * i + 1]; }
    printf("%-40s issue %7.0f cyc/tile  total %7.0f cyc/tile  kernel %.3f ms  -> %.1f GB/s (%s)\n", name, si / (grid * 3) / iters, st / (grid * 3) / iters, ms,
           (double)grid * iters * 33 * 640 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int tiles = 100000;                  // 2 GB of wave: DRAM resident
    float* wave; cudaMalloc(&wave, (size_t)(tiles + 2) * 32 * 160 * 4); cudaMemset(wave, 0, (size_t)(tiles + 2) * 32 * 160 * 4);
    long long* out; cudaMalloc(&out, 148 * 5 * 3 * 2 * 8);
    run<0>(wave, out, tiles, "cp.async8 slot-major (32 lines)");
    run<1>(wave, out, tiles, "cp.async8 frame-major (pitch 161)");
    run<2>(wave, out, tiles, "cp.async8 contiguous raw");
    run<4>(wave, out, tiles, "cp.async16 contiguous raw");
    run<3>(wave, out, tiles, "LDG.64 + 2xSTS.64 slot-major, 1 batch");
    run<5>(wave, out, tiles, "cp.async4 x2 slot-major");
    return 0;
}
