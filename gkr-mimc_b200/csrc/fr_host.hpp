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

// Montgomery product, "no-carry" CIOS (valid because the top limb of q has spare bits: q < 2^254), written with MULX and
// two interleaved carry chains (ADCX / ADOX): one row of x*y[i], then one reduction row, four times.  The transcript is a
// chain of ~2 500 DEPENDENT products per sumcheck round, so what matters here is latency, and gcc's code for the
// portable form (single ADC chain, spills) is ~1.3x slower.  After every row the running value is < x + q, hence any
// x < 2q (and any y < 2^256) is a legal input and the unreduced result is < x*y/2^256 + q.
static const uint64_t ASM_Q[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t ASM_QINV = 0xc2e1f593efffffffULL;

#define GKR_MAC_ROW0                   \
    "movq 0(%[y]), %%rdx\n\t"          \
    "mulx %[x0], %[t0], %[t1]\n\t"     \
    "mulx %[x1], %%rax, %[t2]\n\t"     \
    "addq %%rax, %[t1]\n\t"            \
    "mulx %[x2], %%rax, %[t3]\n\t"     \
    "adcq %%rax, %[t2]\n\t"            \
    "mulx %[x3], %%rax, %[A]\n\t"      \
    "adcq %%rax, %[t3]\n\t"            \
    "adcq $0, %[A]\n\t"
#define GKR_MAC_ROW(off)               \
    "xorl %%eax, %%eax\n\t"            \
    "movq " #off "(%[y]), %%rdx\n\t"   \
    "mulx %[x0], %%rax, %[A]\n\t"      \
    "adox %%rax, %[t0]\n\t"            \
    "adcx %[A], %[t1]\n\t"             \
    "mulx %[x1], %%rax, %[A]\n\t"      \
    "adox %%rax, %[t1]\n\t"            \
    "adcx %[A], %[t2]\n\t"             \
    "mulx %[x2], %%rax, %[A]\n\t"      \
    "adox %%rax, %[t2]\n\t"            \
    "adcx %[A], %[t3]\n\t"             \
    "mulx %[x3], %%rax, %[A]\n\t"      \
    "adox %%rax, %[t3]\n\t"            \
    "movl $0, %%eax\n\t"               \
    "adcx %%rax, %[A]\n\t"             \
    "adox %%rax, %[A]\n\t"
#define GKR_RED_ROW                    \
    "movq %[t0], %%rdx\n\t"            \
    "imulq %[qinv], %%rdx\n\t"         \
    "xorl %%eax, %%eax\n\t"            \
    "mulx %[q0], %%rax, %[H]\n\t"      \
    "adcx %[t0], %%rax\n\t"            \
    "movq %[H], %[t0]\n\t"             \
    "adcx %[t1], %[t0]\n\t"            \
    "mulx %[q1], %%rax, %[t1]\n\t"     \
    "adox %%rax, %[t0]\n\t"            \
    "adcx %[t2], %[t1]\n\t"            \
    "mulx %[q2], %%rax, %[t2]\n\t"     \
    "adox %%rax, %[t1]\n\t"            \
    "adcx %[t3], %[t2]\n\t"            \
    "mulx %[q3], %%rax, %[t3]\n\t"     \
    "adox %%rax, %[t2]\n\t"            \
    "movl $0, %%eax\n\t"               \
    "adcx %%rax, %[t3]\n\t"            \
    "adox %[A], %[t3]\n\t"

// x*y*2^-256 mod q + {0, q}: in [0, 2q) for x, y < 2q (not canonicalised)
static inline __attribute__((always_inline)) Fr mul_lazy(const Fr& x, const Fr& y) {
    uint64_t t0, t1, t2, t3, A, H;
    asm(GKR_MAC_ROW0 GKR_RED_ROW GKR_MAC_ROW(8) GKR_RED_ROW GKR_MAC_ROW(16) GKR_RED_ROW GKR_MAC_ROW(24) GKR_RED_ROW
        : [t0] "=&r"(t0), [t1] "=&r"(t1), [t2] "=&r"(t2), [t3] "=&r"(t3), [A] "=&r"(A), [H] "=&r"(H)
        : [x0] "r"(x.l[0]), [x1] "r"(x.l[1]), [x2] "r"(x.l[2]), [x3] "r"(x.l[3]), [y] "r"(y.l), [q0] "m"(ASM_Q[0]), [q1] "m"(ASM_Q[1]),
          [q2] "m"(ASM_Q[2]), [q3] "m"(ASM_Q[3]), [qinv] "m"(ASM_QINV), "m"(y)
        : "rax", "rdx", "cc");
    return Fr{{t0, t1, t2, t3}};
}
#undef GKR_MAC_ROW0
#undef GKR_MAC_ROW
#undef GKR_RED_ROW
// fr.Element.Mul: canonical inputs, canonical output
static inline __attribute__((always_inline)) Fr mul(const Fr& x, const Fr& y) {
    const Fr t = mul_lazy(x, y);
    return reduce_once(t.l[0], t.l[1], t.l[2], t.l[3]);
}
// x + y without reduction (caller tracks the bound)
static inline __attribute__((always_inline)) Fr add_lazy(const Fr& x, const Fr& y) {
    ull t0, t1, t2, t3;
    unsigned char c = _addcarry_u64(0, x.l[0], y.l[0], &t0);
    c = _addcarry_u64(c, x.l[1], y.l[1], &t1);
    c = _addcarry_u64(c, x.l[2], y.l[2], &t2);
    (void)_addcarry_u64(c, x.l[3], y.l[3], &t3);
    return Fr{{t0, t1, t2, t3}};
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
