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

