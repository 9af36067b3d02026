#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
// Do DFMA and IMAD.WIDE overlap?  Each thread runs WI chains of IMAD.WIDE and WD chains of DFMA interleaved in one loop.
template <int WI, int WD>
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
                if (k < WI) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[k]) : "r"((uint32_t)c[n % (WI ? WI : 1)]), "r"(y));
                if (k < WD) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[k]) : "d"(d[n % (WD ? WD : 1)]), "d"(dy));
            }
        }
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= c[i] ^ (uint64_t)__double_as_longlong(d[i]);
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <typename F> float run(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) { cudaEventRecord(e0); f(rep); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms; }
    return best;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int grid = p.multiProcessorCount * 8, block = 256, iters = 2000; void* d; cudaMalloc(&d, (size_t)grid * block * 8);
    const double per = (double)grid * block * iters * 8.0;
#define RUN(WI, WD) { float ms = run([&](int r) { k<WI, WD><<<grid, block>>>((uint64_t*)d, iters, 7 + r); }); \
    printf("IMAD.WIDE x%d + DFMA x%d per step: %.3f ms -> IMAD.WIDE %.2f T/s, DFMA %.2f T/s\n", WI, WD, ms, per * WI / (ms * 1e-3) / 1e12, per * WD / (ms * 1e-3) / 1e12); }
    RUN(8, 0) RUN(0, 8) RUN(8, 8) RUN(4, 8) RUN(8, 4) RUN(4, 4) RUN(2, 8)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
