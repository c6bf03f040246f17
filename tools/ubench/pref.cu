// Micro-benchmark: cost of pulling a tile's mask rows (2 x 161 rows x 128 B, row stride 4004 B) to L2
//  mode 0: prefetch.global.L2, two per row (both ends)   mode 1: cp.async.bulk.prefetch.L2, one 144-B piece per row
//  mode 2: nothing (baseline loop overhead)
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
template <int MODE>
__global__ void __launch_bounds__(96, 5) k(const float* __restrict__ m, long long* out, int iters, int tmax, int nutt) {
    long long t_issue = 0;
    const int tiles_per = (tmax + 31) / 32;
    for (int it = 0; it < iters; ++it) {
        const int tile = (blockIdx.x + it * gridDim.x) % (tiles_per * nutt);
        const int n = tile / tiles_per, t0 = (tile % tiles_per) * 32;
        const long long c0 = clock64();
        for (int which = 0; which < 2; ++which) {
            const float* base = m + ((long long)(which * nutt + n) * 161) * tmax + t0;
            for (int f = threadIdx.x; f < 161; f += 96) {
                const float* p = base + (long long)f * tmax;
                if (MODE == 0) {
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(p + 31));
                } else if (MODE == 1) {
                    const unsigned long long a = (unsigned long long)p & ~15ull;
                    const unsigned sz = (unsigned)((((unsigned long long)(p + 32) + 15ull) & ~15ull) - a);
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a), "r"(sz) : "memory");
                }
            }
        }
        const long long c1 = clock64();
        t_issue += c1 - c0;
        // some ALU work between tiles (stand-in for the FFT): ~2000 cycles
        float x = (float)tile;
        for (int i = 0; i < 500; ++i) x = x * 1.0001f + 0.5f;
        if (x == 1.2345f) out[0] = 0;
    }
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 3 + (threadIdx.x >> 5)] = t_issue;
}
template <int MODE> void run(const float* m, long long* out, int tmax, int nutt, const char* name) {
    const int grid = 148 * 5, iters = 10;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 96>>>(m, out, 1, tmax, nutt);
    cudaEventRecord(e0);
    k<MODE><<<grid, 96>>>(m, out, iters, tmax, nutt);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid * 3);
    cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost);
    double si = 0; for (auto v : h) si += v;
    printf("%-40s issue %7.0f cyc/tile/warp  kernel %.3f ms -> %.1f GB/s of mask bytes (%s)\n", name, si / (grid * 3) / iters, ms,
           (double)grid * iters * 322 * 128 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int tmax = 1001, nutt = 256;
    const size_t bytes = (size_t)2 * nutt * 161 * tmax * 4 + 4096;
    float* m; cudaMalloc(&m, bytes); cudaMemset(m, 1, bytes);
    long long* out; cudaMalloc(&out, 148 * 5 * 3 * 8);
    run<2>(m, out, tmax, nutt, "nothing");
    run<0>(m, out, tmax, nutt, "prefetch.global.L2 x2 per row");
    run<1>(m, out, tmax, nutt, "cp.async.bulk.prefetch.L2 144 B per row");
    run<0>(m, out, tmax, nutt, "prefetch.global.L2 x2 per row");
    run<1>(m, out, tmax, nutt, "cp.async.bulk.prefetch.L2 144 B per row");
    return 0;
}
