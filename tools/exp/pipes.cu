#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
// Honest pipe-rate microbenchmarks: every multiply takes an operand produced by the previous step of a NEIGHBOUR
// chain, so nothing is loop-invariant and ptxas cannot strength-reduce the multiplies away.
template <int V>
__global__ void __launch_bounds__(256) k(uint64_t* out, int iters, uint32_t seed) {
    uint64_t c[8];
    double d[8];
    uint32_t y = seed | 1;
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i] = seed * (i + 1) + threadIdx.x; d[i] = 1.0 + 1e-9 * (threadIdx.x + i); }
    const double dy = 1.0 + 1e-12 * seed;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int n = (k + 1) & 7;
                if (V == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[k]) : "r"((uint32_t)c[n]), "r"(y));
                if (V == 1) { uint32_t lo = (uint32_t)c[k]; asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"((uint32_t)c[n]), "r"(y)); c[k] = lo; }
                if (V == 2) { uint32_t lo = (uint32_t)c[k]; asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"((uint32_t)c[n]), "r"(y)); c[k] = lo; }
                if (V == 3) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[k]) : "d"(d[n]), "d"(dy));
                if (V == 4) { uint32_t lo = (uint32_t)c[k], hi = (uint32_t)(c[k] >> 32);
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"((uint32_t)c[n]), "r"(y)); c[k] = lo | ((uint64_t)hi << 32); }
            }
        }
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= c[i] ^ (uint64_t)__double_as_longlong(d[i]);
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
    const char* names[] = {"IMAD.WIDE.U32 (64-bit acc)", "IMAD.LO (32-bit)", "IMAD.HI (32-bit)", "DFMA", "IMAD.WIDE via lo.cc/hi.c pair"};
    printf("%-32s %.2f T/s\n", names[0], run([&](int r) { k<0><<<grid, block>>>((uint64_t*)d, iters, 7 + r); }, grid, block, iters));
    printf("%-32s %.2f T/s\n", names[1], run([&](int r) { k<1><<<grid, block>>>((uint64_t*)d, iters, 7 + r); }, grid, block, iters));
    printf("%-32s %.2f T/s\n", names[2], run([&](int r) { k<2><<<grid, block>>>((uint64_t*)d, iters, 7 + r); }, grid, block, iters));
    printf("%-32s %.2f T/s\n", names[3], run([&](int r) { k<3><<<grid, block>>>((uint64_t*)d, iters, 7 + r); }, grid, block, iters));
    printf("%-32s %.2f T/s\n", names[4], run([&](int r) { k<4><<<grid, block>>>((uint64_t*)d, iters, 7 + r); }, grid, block, iters));
    printf("%s (per SM per clk at 1.965 GHz: divide T/s by %.3f)\n", cudaGetErrorString(cudaGetLastError()), p.multiProcessorCount * 1.965e-3);
}
