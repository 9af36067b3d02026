// Experiment: do integer-pipe (IMAD.WIDE) Montgomery products and FP64-pipe products overlap when different warps of the
// same SM run them?  MODE 0: all warps integer, 1: all warps FP64, 2: warp parity decides (odd warps FP64).
#include "fp64mul_core.cuh"
#include "../../gkr-mimc_b200/csrc/fr_device.cuh"
using namespace gkr;

template <int NCH>
__device__ __forceinline__ void run_int(gkr::FrRaw* out, int iters, ull seed) {
    Fr x[NCH], y;
#pragma unroll
    for (int i = 0; i < 8; i++) y.v[i] = (uint32_t)(seed * (i + 1) + threadIdx.x * 7 + blockIdx.x);
    y.v[7] &= 0x0fffffff;
#pragma unroll
    for (int c = 0; c < NCH; c++) { x[c] = y; x[c].v[0] += c + 1; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < NCH; c++) x[c] = fr_mul(x[c], c == 0 ? y : x[c - 1]);
    }
    Fr r = x[0];
#pragma unroll
    for (int c = 1; c < NCH; c++)
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] ^= x[c].v[i];
    fr_store(out + (size_t)blockIdx.x * blockDim.x + threadIdx.x, r);
}
template <int NCH>
__device__ __forceinline__ void run_fp(gkr::FrRaw* out, int iters, ull seed) {
    double q[5];
#pragma unroll
    for (int i = 0; i < 5; i++) q[i] = c_qd[i];
    const ull qinv = c_qinv52;
    FrD x[NCH], y;
    ::FrRaw s;
    s.l[0] = seed * (threadIdx.x + 1); s.l[1] = seed ^ blockIdx.x; s.l[2] = seed + 77 * threadIdx.x; s.l[3] = 0x0123456789abcdefULL;
    y = load52(s);
#pragma unroll
    for (int c = 0; c < NCH; c++) { s.l[0] += 0x9e3779b97f4a7c15ULL; x[c] = load52(s); }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < NCH; c++) x[c] = mul52<false>(x[c], c == 0 ? y : x[c - 1], q, qinv);
    }
    ::FrRaw o = store52(x[0]);
#pragma unroll
    for (int c = 1; c < NCH; c++) { ::FrRaw t = store52(x[c]); o.l[0] ^= t.l[0]; o.l[1] ^= t.l[1]; o.l[2] ^= t.l[2]; o.l[3] ^= t.l[3]; }
    gkr::FrRaw* p = out + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    p->l[0] = o.l[0]; p->l[1] = o.l[1]; p->l[2] = o.l[2]; p->l[3] = o.l[3];
}

template <int MODE, int NCH>
__global__ void __launch_bounds__(256) k_mix(gkr::FrRaw* out, int iters_int, int iters_fp, ull seed) {
    const int warp = threadIdx.x >> 5;
    const bool fp = MODE == 1 || (MODE == 2 && (warp & 1));
    if (fp) run_fp<NCH>(out, iters_fp, seed);
    else run_int<NCH>(out, iters_int, seed);
}

int main() {
    const ull Q[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    double qd[5];
    ull l[5];
    l[0] = Q[0] & MASK52; l[1] = ((Q[0] >> 52) | (Q[1] << 12)) & MASK52; l[2] = ((Q[1] >> 40) | (Q[2] << 24)) & MASK52;
    l[3] = ((Q[2] >> 28) | (Q[3] << 36)) & MASK52; l[4] = Q[3] >> 16;
    for (int i = 0; i < 5; i++) qd[i] = (double)l[i];
    ull inv = 1;
    for (int i = 0; i < 7; i++) inv *= 2 - Q[0] * inv;
    ull qinv52 = (0ULL - inv) & MASK52;
    cudaMemcpyToSymbol(c_qd, qd, sizeof qd);
    cudaMemcpyToSymbol(c_qinv52, &qinv52, sizeof qinv52);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    gkr::FrRaw* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 16 * 256 * 32);
    auto run = [&](auto kern, int bps, int it_int, int it_fp, double muls_per_block, const char* name) {
        int grid = p.multiProcessorCount * bps, block = 256;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0); kern<<<grid, block>>>(d, it_int, it_fp, 12345 + rep); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
        }
        printf("%-10s %d x 256 thr/SM  iters int %5d fp %5d : %.1f G mul/s (%.3f ms) %s\n", name, bps, it_int, it_fp, grid * muls_per_block / (best * 1e-3) / 1e9, best, cudaGetErrorString(cudaGetLastError()));
    };
    const int N = 2;
    for (int bps : {1, 2, 3}) {
        run(k_mix<0, N>, bps, 2000, 0, 256.0 * 2000 * N, "int only");
        run(k_mix<1, N>, bps, 0, 2000, 256.0 * 2000 * N, "fp64 only");
        for (int fpi : {1000, 1400, 1800, 2200, 2600})
            run(k_mix<2, N>, bps, 2000, fpi, 128.0 * (2000 + fpi) * N, "mixed");
    }
    return 0;
}
