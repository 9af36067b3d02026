// BN254 Fr on the host (4 x 64-bit Montgomery) for the serial parts of the prover: the Fiat-Shamir
// transcript (common/challenge.go:10 -> hash/mimc.go:11-49), Lagrange interpolation
// (poly/lagrange.go:96-111) and a handful of per-layer scalars.  Product code: independent of oracle/.
//
// Replaces gnark-crypto's fr.Element (reference go.mod:7; un-vendored).  Same memory image as Go:
// [4]uint64 little-endian limbs, Montgomery form, canonical.
#pragma once
#include <cstdint>
#include <cstring>
#include <x86intrin.h>

namespace gkr {
namespace host {

struct alignas(32) Fr {
    uint64_t l[4];
};

typedef unsigned __int128 u128;

static constexpr uint64_t Q[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static constexpr uint64_t ONE[4] = {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL};
static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};
static constexpr uint64_t QINV = 0xc2e1f593efffffffULL;

typedef unsigned long long ull;

static inline Fr zero() { return Fr{{0, 0, 0, 0}}; }
static inline Fr one() { return Fr{{ONE[0], ONE[1], ONE[2], ONE[3]}}; }
static inline bool eq(const Fr& a, const Fr& b) { return ((a.l[0] ^ b.l[0]) | (a.l[1] ^ b.l[1]) | (a.l[2] ^ b.l[2]) | (a.l[3] ^ b.l[3])) == 0; }
static inline bool is_zero(const Fr& a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }

// t (< 2q) -> canonical
static inline __attribute__((always_inline)) Fr reduce_once(uint64_t t0, uint64_t t1, uint64_t t2, uint64_t t3) {
    ull s0, s1, s2, s3;
    unsigned char b = _subborrow_u64(0, t0, Q[0], &s0);
    b = _subborrow_u64(b, t1, Q[1], &s1);
    b = _subborrow_u64(b, t2, Q[2], &s2);
    b = _subborrow_u64(b, t3, Q[3], &s3);
    Fr z;
    z.l[0] = b ? t0 : s0;
    z.l[1] = b ? t1 : s1;
    z.l[2] = b ? t2 : s2;
    z.l[3] = b ? t3 : s3;
    return z;
}

static inline __attribute__((always_inline)) Fr add(const Fr& x, const Fr& y) {
    ull t0, t1, t2, t3;
    unsigned char c = _addcarry_u64(0, x.l[0], y.l[0], &t0);
    c = _addcarry_u64(c, x.l[1], y.l[1], &t1);
    c = _addcarry_u64(c, x.l[2], y.l[2], &t2);
    (void)_addcarry_u64(c, x.l[3], y.l[3], &t3);  // < 2q < 2^255
    return reduce_once(t0, t1, t2, t3);
}
static inline __attribute__((always_inline)) Fr sub(const Fr& x, const Fr& y) {
    ull t0, t1, t2, t3;
    unsigned char b = _subborrow_u64(0, x.l[0], y.l[0], &t0);
    b = _subborrow_u64(b, x.l[1], y.l[1], &t1);
    b = _subborrow_u64(b, x.l[2], y.l[2], &t2);
    b = _subborrow_u64(b, x.l[3], y.l[3], &t3);
    const uint64_t m = b ? ~0ULL : 0ULL;
    ull r0, r1, r2, r3;
    unsigned char c = _addcarry_u64(0, t0, Q[0] & m, &r0);
    c = _addcarry_u64(c, t1, Q[1] & m, &r1);
    c = _addcarry_u64(c, t2, Q[2] & m, &r2);
    (void)_addcarry_u64(c, t3, Q[3] & m, &r3);
    return Fr{{r0, r1, r2, r3}};
}
static inline Fr neg(const Fr& x) { return sub(zero(), x); }
static inline Fr dbl(const Fr& x) { return add(x, x); }

// Montgomery product, "no-carry" CIOS (valid because the top limb of q has spare bits: q < 2^254).
static inline __attribute__((always_inline)) Fr mul(const Fr& x, const Fr& y) {
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;
#define GKR_ROW(yi)                                                \
    {                                                              \
        u128 a = (u128)x.l[0] * (yi) + t0;                         \
        uint64_t lo = (uint64_t)a, A = (uint64_t)(a >> 64);        \
        uint64_t m = lo * QINV;                                    \
        u128 c = (u128)m * Q[0] + lo;                              \
        uint64_t C = (uint64_t)(c >> 64);                          \
        a = (u128)x.l[1] * (yi) + t1 + A;                          \
        A = (uint64_t)(a >> 64);                                   \
        c = (u128)m * Q[1] + (uint64_t)a + C;                      \
        t0 = (uint64_t)c;                                          \
        C = (uint64_t)(c >> 64);                                   \
        a = (u128)x.l[2] * (yi) + t2 + A;                          \
        A = (uint64_t)(a >> 64);                                   \
        c = (u128)m * Q[2] + (uint64_t)a + C;                      \
        t1 = (uint64_t)c;                                          \
        C = (uint64_t)(c >> 64);                                   \
        a = (u128)x.l[3] * (yi) + t3 + A;                          \
        A = (uint64_t)(a >> 64);                                   \
        c = (u128)m * Q[3] + (uint64_t)a + C;                      \
        t2 = (uint64_t)c;                                          \
        C = (uint64_t)(c >> 64);                                   \
        t3 = C + A;                                                \
    }
    GKR_ROW(y.l[0])
    GKR_ROW(y.l[1])
    GKR_ROW(y.l[2])
    GKR_ROW(y.l[3])
#undef GKR_ROW
    return reduce_once(t0, t1, t2, t3);
}
static inline __attribute__((always_inline)) Fr sqr(const Fr& x) { return mul(x, x); }

static inline Fr from_u64(uint64_t v) {  // fr.Element.SetUint64
    Fr a{{v, 0, 0, 0}}, r2{{R2[0], R2[1], R2[2], R2[3]}};
    return mul(a, r2);
}
static inline Fr to_mont(const Fr& x) { return mul(x, Fr{{R2[0], R2[1], R2[2], R2[3]}}); }
static inline Fr from_mont(const Fr& x) { return mul(x, Fr{{1, 0, 0, 0}}); }

// fr.Element.Inverse (0 -> 0): x^(q-2)
static inline Fr inv(const Fr& x) {
    uint64_t e[4] = {Q[0] - 2, Q[1], Q[2], Q[3]};
    Fr acc = one(), base = x;
    for (int i = 0; i < 256; i++) {
        if ((e[i / 64] >> (i % 64)) & 1) acc = mul(acc, base);
        base = sqr(base);
    }
    return acc;
}

}  // namespace host
}  // namespace gkr
