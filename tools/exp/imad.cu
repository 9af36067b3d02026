#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
// variants of IMAD.WIDE throughput: V=0 same x,y for all accumulators; V=1 distinct x_k, shared y; V=2 distinct x_k, y_k; V=3: like V=1 but carry-chained (.X)
template <int V>
__global__ void __launch_bounds__(256) k(uint64_t* out, int iters, uint32_t seed) {
    uint32_t x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = seed * (i + 1) + threadIdx.x; y[i] = seed * (i + 11) + blockIdx.x; }
    uint64_t c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = i;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t xx = V == 0 ? x[0] : x[k];
                if (V == 5) { asm volatile("mad.wide.u32 %0, %1, 0x1234567, %0;" : "+l"(c[k]) : "r"(xx)); continue; }
                const uint32_t yy = V == 2 ? y[(k + u) & 7] : (V == 4 ? seed + u : y[u]);
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[k]) : "r"(xx), "r"(yy));
            }
        }
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= c[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// 32-bit IMAD lo / hi
template <int V>
__global__ void __launch_bounds__(256) k32(uint32_t* out, int iters, uint32_t seed) {
    uint32_t x[8], y[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = seed * (i + 1) + threadIdx.x; y[i] = seed * (i + 11) + blockIdx.x; c[i] = i; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (V == 0) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c[k]) : "r"(x[k]), "r"(y[u]));
                else asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(c[k]) : "r"(x[k]), "r"(y[u]));
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= c[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <typename F> double run(F f, int grid, int block, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) { cudaEventRecord(e0); f(rep); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms; }
    return (double)grid * block * iters * 64.0 / (best * 1e-3) / 1e12;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int grid = p.multiProcessorCount * 8, block = 256, iters = 2000; void* d; cudaMalloc(&d, (size_t)grid * block * 8);
    printf("IMAD.WIDE same x,y      : %.2f T/s\n", run([&](int r) { k<0><<<grid, block>>>((uint64_t*)d, iters, 7 + r); }, grid, block, iters));
    printf("IMAD.WIDE distinct x    : %.2f T/s\n", run([&](int r) { k<1><<<grid, block>>>((uint64_t*)d, iters, 7 + r); }, grid, block, iters));
    printf("IMAD.WIDE distinct x,y  : %.2f T/s\n", run([&](int r) { k<2><<<grid, block>>>((uint64_t*)d, iters, 7 + r); }, grid, block, iters));
    printf("IMAD.WIDE x reg, y uniform : %.2f T/s\n", run([&](int r) { k<4><<<grid, block>>>((uint64_t*)d, iters, 7 + r); }, grid, block, iters));
    printf("IMAD.WIDE x reg, y imm     : %.2f T/s\n", run([&](int r) { k<5><<<grid, block>>>((uint64_t*)d, iters, 7 + r); }, grid, block, iters));
    printf("IMAD.LO  distinct x     : %.2f T/s\n", run([&](int r) { k32<0><<<grid, block>>>((uint32_t*)d, iters, 7 + r); }, grid, block, iters));
    printf("IMAD.HI  distinct x     : %.2f T/s\n", run([&](int r) { k32<1><<<grid, block>>>((uint32_t*)d, iters, 7 + r); }, grid, block, iters));
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
