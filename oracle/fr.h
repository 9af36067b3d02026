/* BN254 scalar field (Fr) arithmetic, 4x64 Montgomery form -- ORACLE (test infrastructure).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * link or load anything under oracle/.  The product library never does.
 *
 * Restates the semantics of the third-party dependency
 *   github.com/consensys/gnark-crypto v0.6.1-0.20220110145513-493bb1c180d9, ecc/bn254/fr
 * (reference go.mod:7; source NOT vendored under the reference tree): `Element [4]uint64`,
 * little-endian limbs, value stored as a*2^256 mod q, always fully reduced (< q).
 * Constants: SURVEY.md Appendix A (re-derived in tests/test_oracle.py from q alone).
 */
#ifndef ORC_FR_H
#define ORC_FR_H
#include <stdint.h>
#include <string.h>

typedef struct { uint64_t l[4]; } fr_t;
typedef unsigned __int128 u128;

static const uint64_t FR_Q[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t FR_ONE[4] = {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL};
static const uint64_t FR_R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};
#define FR_QINV 0xc2e1f593efffffffULL /* -q^{-1} mod 2^64 */
static const uint64_t FR_QINV_M = FR_QINV;

static inline int fr_geq_q(const uint64_t t[4]) {
    for (int i = 3; i >= 0; i--) {
        if (t[i] > FR_Q[i]) return 1;
        if (t[i] < FR_Q[i]) return 0;
    }
    return 1;
}
static inline void fr_sub_q(uint64_t t[4]) {
    u128 b = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)t[i] - FR_Q[i] - (uint64_t)b;
        t[i] = (uint64_t)d;
        b = (d >> 64) & 1;
    }
}
/* fr.Element.Add */
static inline void fr_add(fr_t *z, const fr_t *x, const fr_t *y) {
    uint64_t t[4];
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)x->l[i] + y->l[i];
        t[i] = (uint64_t)c;
        c >>= 64;
    } /* q < 2^254 so no carry out of limb 3 */
    if (fr_geq_q(t)) fr_sub_q(t);
    memcpy(z->l, t, 32);
}
/* fr.Element.Sub */
static inline void fr_sub(fr_t *z, const fr_t *x, const fr_t *y) {
    uint64_t t[4];
    u128 b = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)x->l[i] - y->l[i] - (uint64_t)b;
        t[i] = (uint64_t)d;
        b = (d >> 64) & 1;
    }
    if (b) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)t[i] + FR_Q[i];
            t[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    memcpy(z->l, t, 32);
}
/* fr.Element.Mul : Montgomery product x*y*2^-256 mod q.
 * fr_mul_portable: textbook CIOS on unsigned __int128 (kept as the cross-check of the fast form, tests/test_oracle.py). */
static inline void fr_mul_portable(fr_t *z, const fr_t *x, const fr_t *y) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)x->l[j] * y->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * FR_QINV;
        c = (u128)m * FR_Q[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * FR_Q[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || fr_geq_q(t)) fr_sub_q(t);
    memcpy(z->l, t, 32);
}

#if defined(__x86_64__) && defined(__BMI2__) && defined(__ADX__) && !defined(ORC_FR_PORTABLE)
/* The multiplier the reference actually runs is gnark-crypto's amd64 assembly (element_ops_amd64.s / mul_amd64.s of the pinned
 * module, not vendored): MULX with two carry chains (ADCX / ADOX), the "no-carry" CIOS of gnark's modular-multiplication note
 * (valid because q's top limb has spare bits).  This is that published algorithm written out for gcc: per row i,
 *    (A, t) += x * y[i];   m = t0 * qInvNeg;   (C, t) = (t + m * q) >> 64;   t3 = A + C
 * with the low and high halves of the MULX products on separate flags (OF / CF), so the CPU baseline is timed with the same
 * class of multiplier as the Go binary.  BASELINE.md section 3 reports its ns per product. */
#define ORC_ROW_XY(off)                                                                                                      \
    "xorl %%eax, %%eax\n\t"                                                                                                  \
    "movq " #off "(%[y]), %%rdx\n\t"                                                                                         \
    "mulx 0(%[x]), %%rax, %%rbx\n\t"                                                                                         \
    "adox %%rax, %[t0]\n\t"                                                                                                  \
    "adcx %%rbx, %[t1]\n\t"                                                                                                  \
    "mulx 8(%[x]), %%rax, %%rbx\n\t"                                                                                         \
    "adox %%rax, %[t1]\n\t"                                                                                                  \
    "adcx %%rbx, %[t2]\n\t"                                                                                                  \
    "mulx 16(%[x]), %%rax, %%rbx\n\t"                                                                                        \
    "adox %%rax, %[t2]\n\t"                                                                                                  \
    "adcx %%rbx, %[t3]\n\t"                                                                                                  \
    "mulx 24(%[x]), %%rax, %[A]\n\t"                                                                                         \
    "adox %%rax, %[t3]\n\t"                                                                                                  \
    "movl $0, %%eax\n\t"                                                                                                     \
    "adcx %%rax, %[A]\n\t"                                                                                                   \
    "adox %%rax, %[A]\n\t"
#define ORC_ROW_MQ                                                                                                           \
    "movq %[t0], %%rdx\n\t"                                                                                                  \
    "imulq %[qinv], %%rdx\n\t"                                                                                               \
    "xorl %%eax, %%eax\n\t"                                                                                                  \
    "mulx %[q0], %%rax, %%rbx\n\t"                                                                                           \
    "adcx %[t0], %%rax\n\t" /* low limb cancels; CF = its carry */                                                         \
    "movq %%rbx, %[t0]\n\t"                                                                                                  \
    "adcx %[t1], %[t0]\n\t"                                                                                                  \
    "mulx %[q1], %%rax, %[t1]\n\t"                                                                                           \
    "adox %%rax, %[t0]\n\t"                                                                                                  \
    "adcx %[t2], %[t1]\n\t"                                                                                                  \
    "mulx %[q2], %%rax, %[t2]\n\t"                                                                                           \
    "adox %%rax, %[t1]\n\t"                                                                                                  \
    "adcx %[t3], %[t2]\n\t"                                                                                                  \
    "mulx %[q3], %%rax, %[t3]\n\t"                                                                                           \
    "adox %%rax, %[t2]\n\t"                                                                                                  \
    "movl $0, %%eax\n\t"                                                                                                     \
    "adcx %%rax, %[t3]\n\t"                                                                                                  \
    "adox %[A], %[t3]\n\t"
static inline void fr_mul(fr_t *z, const fr_t *x, const fr_t *y) {
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, A;
    __asm__(ORC_ROW_XY(0) ORC_ROW_MQ ORC_ROW_XY(8) ORC_ROW_MQ ORC_ROW_XY(16) ORC_ROW_MQ ORC_ROW_XY(24) ORC_ROW_MQ
            : [t0] "+&r"(t0), [t1] "+&r"(t1), [t2] "+&r"(t2), [t3] "+&r"(t3), [A] "=&r"(A)
            : [x] "r"(x->l), [y] "r"(y->l), [q0] "m"(FR_Q[0]), [q1] "m"(FR_Q[1]), [q2] "m"(FR_Q[2]), [q3] "m"(FR_Q[3]), [qinv] "m"(FR_QINV_M),
              "m"(*x), "m"(*y)
            : "rax", "rbx", "rdx", "cc");
    uint64_t t[4] = {t0, t1, t2, t3}; /* < 2q: one conditional subtraction (fr.Element.Mul's final step) */
    if (fr_geq_q(t)) fr_sub_q(t);
    memcpy(z->l, t, 32);
}
#undef ORC_ROW_XY
#undef ORC_ROW_MQ
#define ORC_FR_MUL_KIND "mulx/adcx/adox no-carry CIOS (gnark-crypto amd64 class)"
#else
static inline void fr_mul(fr_t *z, const fr_t *x, const fr_t *y) { fr_mul_portable(z, x, y); }
#define ORC_FR_MUL_KIND "portable unsigned __int128 CIOS"
#endif
static inline void fr_sqr(fr_t *z, const fr_t *x) { fr_mul(z, x, x); }
static inline void fr_set_one(fr_t *z) { memcpy(z->l, FR_ONE, 32); }
static inline void fr_set_zero(fr_t *z) { memset(z->l, 0, 32); }
static inline int fr_eq(const fr_t *x, const fr_t *y) { return memcmp(x->l, y->l, 32) == 0; }
static inline int fr_is_zero(const fr_t *x) { return (x->l[0] | x->l[1] | x->l[2] | x->l[3]) == 0; }
/* fr.Element.SetUint64 : {v,0,0,0} * R^2 */
static inline void fr_set_u64(fr_t *z, uint64_t v) {
    fr_t a = {{v, 0, 0, 0}}, r2;
    memcpy(r2.l, FR_R2, 32);
    fr_mul(z, &a, &r2);
}
/* regular (canonical, < q) limbs -> Montgomery */
static inline void fr_to_mont(fr_t *z, const fr_t *x) {
    fr_t r2;
    memcpy(r2.l, FR_R2, 32);
    fr_mul(z, x, &r2);
}
/* fr.Element.FromMont / ToBigIntRegular */
static inline void fr_from_mont(fr_t *z, const fr_t *x) {
    fr_t one = {{1, 0, 0, 0}};
    fr_mul(z, x, &one);
}
static inline void fr_neg(fr_t *z, const fr_t *x) {
    fr_t zero = {{0, 0, 0, 0}};
    fr_sub(z, &zero, x);
}
static inline void fr_dbl(fr_t *z, const fr_t *x) { fr_add(z, x, x); }
/* fr.Element.Inverse (0 -> 0) via Fermat x^(q-2) */
static inline void fr_inv(fr_t *z, const fr_t *x) {
    uint64_t e[4];
    memcpy(e, FR_Q, 32);
    e[0] -= 2;
    fr_t acc, base = *x;
    fr_set_one(&acc);
    for (int i = 0; i < 256; i++) {
        if ((e[i / 64] >> (i % 64)) & 1) fr_mul(&acc, &acc, &base);
        fr_sqr(&base, &base);
    }
    *z = acc;
}
#endif
