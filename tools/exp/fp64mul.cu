// Experiment: BN254 Fr Montgomery product (R = 2^256, same representation as the Go memory image) on the FP64 pipe.
// 5 limbs x 52 bits held as doubles; every 52x52 product is split exactly with two DFMA.RZ and one DADD
//   hi = fma_rz(a, b, 2^104)            = 2^104 + floor(ab / 2^52) * 2^52
//   lo = fma_rz(a, b, (2^104+2^52) - hi) = 2^52 + (ab mod 2^52)
// and the BIT PATTERNS of hi / lo are summed as 64-bit integers per column (the exponent fields add up to a
// compile-time constant that the accumulators are pre-loaded with).  Reduction rows are 52,52,52,52,48 bits wide.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long ull;
struct FrD { double v[5]; };     // limbs < 2^52, value = sum v[i] * 2^(52 i)
struct FrRaw { ull l[4]; };

#define MASK52 0xfffffffffffffULL
#define HI_C 0x4670000000000000ULL  // bit pattern of 2^104
#define LO_C 0x4330000000000000ULL  // bit pattern of 2^52

__constant__ double c_qd[5];
__constant__ ull c_qinv52;  // -q^-1 mod 2^52

__host__ __device__ constexpr int npairs(int k) { return k < 0 ? 0 : (k <= 4 ? k + 1 : (k <= 8 ? 9 - k : 0)); }
// what column k receives over the whole product: 2*npairs(k) lo patterns and 2*npairs(k-1) hi patterns
__host__ __device__ constexpr ull col_init(int k) { return 0ULL - (2ULL * npairs(k) * LO_C + 2ULL * npairs(k - 1) * HI_C); }

__device__ __forceinline__ double u2d(ull x) { return __longlong_as_double((long long)(x | LO_C)) - 4503599627370496.0; }

template <bool CANON>
__device__ __forceinline__ FrD mul52(const FrD& a, const FrD& b, const double (&q)[5], ull qinv) {
    const double C1 = 20282409603651670423947251286016.0;                 // 2^104
    const double C2 = 20282409603651670423947251286016.0 + 4503599627370496.0;  // 2^104 + 2^52 (exact)
    ull acc[10];
#pragma unroll
    for (int k = 0; k < 10; k++) acc[k] = col_init(k);
#pragma unroll
    for (int i = 0; i < 5; i++) {
        {
            ull hp = 0;
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const double hi = __fma_rz(a.v[j], b.v[i], C1);
                const double lo = __fma_rz(a.v[j], b.v[i], C2 - hi);
                acc[i + j] = acc[i + j] + (ull)__double_as_longlong(lo) + hp;
                hp = (ull)__double_as_longlong(hi);
            }
            acc[i + 5] += hp;
        }
        const ull m = (acc[i] * qinv) & (i == 4 ? 0xffffffffffffULL : MASK52);
        const double md = u2d(m);
        {
            ull hp = 0;
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const double hi = __fma_rz(md, c_qd[j], C1);
                const double lo = __fma_rz(md, c_qd[j], C2 - hi);
                acc[i + j] = acc[i + j] + (ull)__double_as_longlong(lo) + hp;
                hp = (ull)__double_as_longlong(hi);
            }
            acc[i + 5] += hp;
        }
        if (i < 4) acc[i + 1] += acc[i] >> 52;
    }
    // value = (acc[4] >> 48) + sum_{t>=0} acc[5+t] * 2^(4 + 52 t)
    ull c = acc[4] >> 48;
    ull r[5];
#pragma unroll
    for (int t = 0; t < 5; t++) {
        const ull v = c + (acc[5 + t] << 4);
        r[t] = v & MASK52;
        c = v >> 52;
    }
    FrD o;
#pragma unroll
    for (int t = 0; t < 5; t++) o.v[t] = u2d(r[t]);
    (void)CANON;
    return o;
}

__device__ __forceinline__ FrD load52(const FrRaw& x) {
    FrD o;
    o.v[0] = u2d(x.l[0] & MASK52);
    o.v[1] = u2d(((x.l[0] >> 52) | (x.l[1] << 12)) & MASK52);
    o.v[2] = u2d(((x.l[1] >> 40) | (x.l[2] << 24)) & MASK52);
    o.v[3] = u2d(((x.l[2] >> 28) | (x.l[3] << 36)) & MASK52);
    o.v[4] = u2d(x.l[3] >> 16);
    return o;
}
__device__ __forceinline__ FrRaw store52(const FrD& a) {
    ull r[5];
#pragma unroll
    for (int t = 0; t < 5; t++) r[t] = (ull)__double_as_longlong(a.v[t] + 4503599627370496.0) & MASK52;
    FrRaw o;
    o.l[0] = r[0] | (r[1] << 52);
    o.l[1] = (r[1] >> 12) | (r[2] << 40);
    o.l[2] = (r[2] >> 24) | (r[3] << 28);
    o.l[3] = (r[3] >> 36) | (r[4] << 16);
    return o;
}

__global__ void k_check(const FrRaw* a, const FrRaw* b, FrRaw* out, int n) {
    double q[5];
#pragma unroll
    for (int i = 0; i < 5; i++) q[i] = c_qd[i];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    out[t] = store52(mul52<false>(load52(a[t]), load52(b[t]), q, c_qinv52));
}

template <int NCH>
__global__ void __launch_bounds__(128) k_bench(FrRaw* out, int iters, ull seed) {
    double q[5];
#pragma unroll
    for (int i = 0; i < 5; i++) q[i] = c_qd[i];
    const ull qinv = c_qinv52;
    FrD x[NCH], y;
    FrRaw s;
    s.l[0] = seed * (threadIdx.x + 1); s.l[1] = seed ^ blockIdx.x; s.l[2] = seed + 77 * threadIdx.x; s.l[3] = 0x0123456789abcdefULL;
    y = load52(s);
#pragma unroll
    for (int c = 0; c < NCH; c++) { s.l[0] += 0x9e3779b97f4a7c15ULL; x[c] = load52(s); }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < NCH; c++) x[c] = mul52<false>(x[c], c == 0 ? y : x[c - 1], q, qinv);
    }
    FrRaw o = store52(x[0]);
#pragma unroll
    for (int c = 1; c < NCH; c++) { FrRaw t = store52(x[c]); o.l[0] ^= t.l[0]; o.l[1] ^= t.l[1]; o.l[2] ^= t.l[2]; o.l[3] ^= t.l[3]; }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = o;
}

// ---- host reference: plain Montgomery with __int128 ------------------------------------------------
typedef unsigned __int128 u128;
static const ull Q[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const ull QINV64 = 0xc2e1f593efffffffULL;
static void host_mont(const ull* a, const ull* b, ull* r /* NOT canonical: < ab/R + q */) {
    ull t[9] = {0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a[j] * b[i] + t[i + j]; t[i + j] = (ull)c; c >>= 64; }
        for (int k = i + 4; c; k++) { c += t[k]; t[k] = (ull)c; c >>= 64; }
    }
    for (int i = 0; i < 4; i++) {
        const ull m = t[i] * QINV64;
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)m * Q[j] + t[i + j]; t[i + j] = (ull)c; c >>= 64; }
        for (int k = i + 4; c; k++) { c += t[k]; t[k] = (ull)c; c >>= 64; }
    }
    for (int i = 0; i < 4; i++) r[i] = t[4 + i];
}
static ull rnd_state = 88172645463325252ULL;
static ull rnd() { rnd_state ^= rnd_state << 13; rnd_state ^= rnd_state >> 7; rnd_state ^= rnd_state << 17; return rnd_state; }

int main(int argc, char** argv) {
    // q in 52-bit limbs, -q^-1 mod 2^52
    double qd[5];
    {
        ull l[5];
        l[0] = Q[0] & MASK52; l[1] = ((Q[0] >> 52) | (Q[1] << 12)) & MASK52; l[2] = ((Q[1] >> 40) | (Q[2] << 24)) & MASK52;
        l[3] = ((Q[2] >> 28) | (Q[3] << 36)) & MASK52; l[4] = Q[3] >> 16;
        for (int i = 0; i < 5; i++) qd[i] = (double)l[i];
    }
    ull inv = 1;
    for (int i = 0; i < 7; i++) inv *= 2 - Q[0] * inv;
    ull qinv52 = (0ULL - inv) & MASK52;
    cudaMemcpyToSymbol(c_qd, qd, sizeof qd);
    cudaMemcpyToSymbol(c_qinv52, &qinv52, sizeof qinv52);

    // correctness: random < 2^254 and edge values
    const int n = 1 << 16;
    FrRaw *ha = (FrRaw*)malloc(n * 32), *hb = (FrRaw*)malloc(n * 32), *hr = (FrRaw*)malloc(n * 32);
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < 4; k++) { ha[i].l[k] = rnd(); hb[i].l[k] = rnd(); }
        ha[i].l[3] &= 0x3fffffffffffffffULL; hb[i].l[3] &= 0x3fffffffffffffffULL;
        if (i < 8) for (int k = 0; k < 4; k++) ha[i].l[k] = (i & 1) ? Q[k] - (k == 0) : 0, hb[i].l[k] = (i & 2) ? Q[k] - (k == 0) : (i & 4 ? ~0ULL >> (k == 3 ? 2 : 0) : 1);
        if (i >= 8 && i < 16) for (int k = 0; k < 4; k++) ha[i].l[k] = ~0ULL >> (k == 3 ? 2 : 0);
    }
    FrRaw *da, *db, *dr;
    cudaMalloc(&da, n * 32); cudaMalloc(&db, n * 32); cudaMalloc(&dr, n * 32);
    cudaMemcpy(da, ha, n * 32, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, n * 32, cudaMemcpyHostToDevice);
    k_check<<<n / 128, 128>>>(da, db, dr, n);
    cudaMemcpy(hr, dr, n * 32, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < n; i++) {
        ull e[4];
        host_mont(ha[i].l, hb[i].l, e);
        if (e[0] != hr[i].l[0] || e[1] != hr[i].l[1] || e[2] != hr[i].l[2] || e[3] != hr[i].l[3]) {
            if (bad < 4) printf("MISMATCH %d: got %016llx %016llx %016llx %016llx want %016llx %016llx %016llx %016llx\n", i, hr[i].l[3], hr[i].l[2], hr[i].l[1], hr[i].l[0], e[3], e[2], e[1], e[0]);
            bad++;
        }
    }
    printf("check: %d / %d mismatches (%s)\n", bad, n, cudaGetErrorString(cudaGetLastError()));

    int iters = argc > 1 ? atoi(argv[1]) : 2000;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    FrRaw* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 16 * 128 * 32);
    auto run = [&](auto kern, int nch, int bps, const char* name) {
        int grid = p.multiProcessorCount * bps, block = 128;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0); kern<<<grid, block>>>(d, iters, 12345 + rep); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
        }
        printf("%s: %d chains/thread, %d x 128 threads/SM: %.1f G mul/s (%.3f ms) %s\n", name, nch, bps, (double)grid * block * iters * nch / (best * 1e-3) / 1e9, best, cudaGetErrorString(cudaGetLastError()));
    };
    for (int bps : {2, 4, 6, 8}) run(k_bench<1>, 1, bps, "mul52");
    for (int bps : {2, 4, 6, 8}) run(k_bench<2>, 2, bps, "mul52");
    for (int bps : {2, 4}) run(k_bench<4>, 4, bps, "mul52");
    return 0;
}
