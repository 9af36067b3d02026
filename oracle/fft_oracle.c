/* CPU restatement of the FFT half of the Groth16 prover -- ORACLE (test infrastructure).
 *
 * Only tests/ and bench tools' cpu_baseline legs may load this; the product libraries never do.
 *
 * What it follows:
 *   prover/gadget/prove.go:310-366   computeH(a, b, c, domain): 3 x FFTInverse(DIF, 0), 3 x FFT(DIT, 1), h = (a o b - c) / -2,
 *                                    FFTInverse(DIF, 1), FromMont
 *   pkg/gnark/notinternal/backend/bn254/groth16/setup.go:98   domain = fft.NewDomain(len(r1cs.Constraints), 1, true)
 * The transform lives in the un-vendored github.com/consensys/gnark-crypto v0.6.1-0.20220110145513-493bb1c180d9 (reference go.mod:7),
 * package ecc/bn254/fr/fft.  Its published definitions (see oracle/pyref_fft.py for the formulas) are restated with the TEXTBOOK
 * algorithm: explicit bit-reversal permutation, then iterative decimation-in-time butterflies with the twiddle advanced by one
 * multiplication per butterfly; every FFT/FFTInverse variant is expressed through that one natural-order transform plus
 * permutations and scalings.  The product's kernels (csrc/ec/ntt.cuh: in-register radix-8 passes over a twiddle table, in-place
 * DIF/DIT without any permutation) share nothing with it.  All results are exact field elements: parity is bit-exact.
 *
 * Pinning: the reference holds no golden vector for this path.  The oracle is pinned by (1) the published 2^28-th root of unity of
 * BN254 Fr, checked to equal 5^((q-1)/2^28) and to have order exactly 2^28, (2) agreement with the O(n^2) definitions in
 * oracle/pyref_fft.py, (3) the meaning of h: for a satisfied system, h is the quotient (A B - C) / (X^n - 1), computed there by
 * polynomial long division with no FFT at all (tests/test_ntt_cpu.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "fr.h"

/* 19103219067921713944291392827692070036145651957329286315305642004821462161904, regular form, little-endian limbs */
static const uint64_t ROOT_OF_UNITY[4] = {0x9bd61b6e725b19f0ULL, 0x402d111e41112ed4ULL, 0x00e0a7eb8ef62abcULL, 0x2a3c09f0a58a7e85ULL};
#define MAX_ORDER_ROOT 28

static void fr_pow2k(fr_t *z, const fr_t *x, unsigned k) { /* x^(2^k) */
    *z = *x;
    for (unsigned i = 0; i < k; i++) fr_sqr(z, z);
}
static unsigned log2u(size_t n) {
    unsigned l = 0;
    while (((size_t)1 << l) < n) l++;
    return l;
}
static size_t revbits(size_t i, unsigned log) {
    size_t r = 0;
    for (unsigned b = 0; b < log; b++) r = (r << 1) | ((i >> b) & 1);
    return r;
}
static void bit_reverse(fr_t *a, size_t n, unsigned log) {
    for (size_t i = 0; i < n; i++) {
        const size_t j = revbits(i, log);
        if (i < j) {
            fr_t t = a[i];
            a[i] = a[j];
            a[j] = t;
        }
    }
}
/* Domain.Generator and Domain.FinerGenerator^coset (depth 1), Montgomery form */
static void domain_roots(unsigned log, int coset, fr_t *w, fr_t *shift) {
    fr_t g, g_reg;
    memcpy(g_reg.l, ROOT_OF_UNITY, 32);
    fr_to_mont(&g, &g_reg);
    fr_pow2k(w, &g, MAX_ORDER_ROOT - log);
    fr_t u;
    fr_pow2k(&u, &g, MAX_ORDER_ROOT - log - 1);
    fr_set_one(shift);
    for (int k = 0; k < coset; k++) fr_mul(shift, shift, &u);
}
/* natural order in, natural order out: a[k] <- sum_j a[j] w^(jk) */
static void ntt_natural(fr_t *a, size_t n, unsigned log, const fr_t *w) {
    bit_reverse(a, n, log);
    for (unsigned s = 1; s <= log; s++) {
        const size_t m = (size_t)1 << s, half = m >> 1;
        fr_t wm;
        fr_pow2k(&wm, w, log - s); /* primitive m-th root */
        for (size_t k = 0; k < n; k += m) {
            fr_t t;
            fr_set_one(&t);
            for (size_t j = 0; j < half; j++) {
                fr_t v, u = a[k + j];
                fr_mul(&v, &a[k + j + half], &t);
                fr_add(&a[k + j], &u, &v);
                fr_sub(&a[k + j + half], &u, &v);
                fr_mul(&t, &t, &wm);
            }
        }
    }
}
static void scale_powers(fr_t *a, size_t n, const fr_t *s, const fr_t *first) { /* a[j] *= first * s^j */
    fr_t t = *first;
    for (size_t j = 0; j < n; j++) {
        fr_mul(&a[j], &a[j], &t);
        fr_mul(&t, &t, s);
    }
}

/* decimation: 0 = DIT (bit-reversed input, natural output), 1 = DIF (natural input, bit-reversed output); coset 0 or 1 */
void orc_fft(uint64_t *data, size_t n, int decimation, int coset) {
    fr_t *a = (fr_t *)data;
    const unsigned log = log2u(n);
    fr_t w, s, one;
    domain_roots(log, coset, &w, &s);
    fr_set_one(&one);
    if (decimation == 0) bit_reverse(a, n, log);
    if (coset) scale_powers(a, n, &s, &one);
    ntt_natural(a, n, log, &w);
    if (decimation == 1) bit_reverse(a, n, log);
}
void orc_fft_inverse(uint64_t *data, size_t n, int decimation, int coset) {
    fr_t *a = (fr_t *)data;
    const unsigned log = log2u(n);
    fr_t w, s, w_inv, s_inv, n_inv, nn;
    domain_roots(log, coset, &w, &s);
    fr_inv(&w_inv, &w);
    fr_inv(&s_inv, &s);
    fr_set_u64(&nn, (uint64_t)n);
    fr_inv(&n_inv, &nn);
    if (decimation == 0) bit_reverse(a, n, log);
    ntt_natural(a, n, log, &w_inv);
    scale_powers(a, n, &s_inv, &n_inv);
    if (decimation == 1) bit_reverse(a, n, log);
}

/* computeH (prove.go:310-366).  a, b, c: n_in Montgomery elements each; h_out: `cardinality` elements in REGULAR form (the last step
 * is FromMont), coefficients in bit-reversed order exactly as the Go code leaves them (pk.G1.Z is stored bit-reversed to match,
 * setup.go:229).  cardinality = next power of two >= the number of constraints.                                                    */
void orc_compute_h(const uint64_t *a_in, const uint64_t *b_in, const uint64_t *c_in, size_t n_in, size_t cardinality, uint64_t *h_out) {
    const size_t n = cardinality;
    fr_t *a = (fr_t *)calloc(n, sizeof(fr_t)), *b = (fr_t *)calloc(n, sizeof(fr_t)), *c = (fr_t *)calloc(n, sizeof(fr_t));
    memcpy(a, a_in, n_in * 32); /* padding with zeros (prove.go:319-323) */
    memcpy(b, b_in, n_in * 32);
    memcpy(c, c_in, n_in * 32);
    orc_fft_inverse((uint64_t *)a, n, 1, 0);
    orc_fft_inverse((uint64_t *)b, n, 1, 0);
    orc_fft_inverse((uint64_t *)c, n, 1, 0);
    orc_fft((uint64_t *)a, n, 0, 1);
    orc_fft((uint64_t *)b, n, 0, 1);
    orc_fft((uint64_t *)c, n, 0, 1);
    fr_t two, minus_two_inv;
    fr_set_u64(&two, 2);
    fr_neg(&two, &two);
    fr_inv(&minus_two_inv, &two);
    for (size_t i = 0; i < n; i++) {
        fr_mul(&a[i], &a[i], &b[i]);
        fr_sub(&a[i], &a[i], &c[i]);
        fr_mul(&a[i], &a[i], &minus_two_inv);
    }
    orc_fft_inverse((uint64_t *)a, n, 1, 1);
    for (size_t i = 0; i < n; i++) fr_from_mont(&a[i], &a[i]);
    memcpy(h_out, a, n * 32);
    free(a);
    free(b);
    free(c);
}

/* Domain constants for the tests: out[0..4) Generator, out[4..8) FinerGenerator, out[8..12) CardinalityInv (Montgomery) */
void orc_fft_domain(size_t cardinality, uint64_t *out) {
    fr_t w, u, nn, n_inv;
    domain_roots(log2u(cardinality), 1, &w, &u);
    fr_set_u64(&nn, (uint64_t)cardinality);
    fr_inv(&n_inv, &nn);
    memcpy(out, w.l, 32);
    memcpy(out + 4, u.l, 32);
    memcpy(out + 8, n_inv.l, 32);
}

/* ---- size-independent check of h (tests at 2^22): evaluate both sides of  H(z) (z^n - 1) = A(z) B(z) - C(z)  at one point ---- */
void orc_fr_mul_elementwise(const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out) {
    for (size_t i = 0; i < n; i++) fr_mul((fr_t *)out + i, (const fr_t *)a + i, (const fr_t *)b + i);
}
/* P(z) for the polynomial whose coefficient rev(i) sits at coef[i] (the order computeH leaves); coef in REGULAR form if
 * `regular`, else Montgomery; z and the result in Montgomery form.  Horner over the natural order.                               */
void orc_eval_poly_bitrev(const uint64_t *coef, size_t n, int regular, const uint64_t *z, uint64_t *out) {
    const unsigned log = log2u(n);
    const fr_t *c = (const fr_t *)coef, *zz = (const fr_t *)z;
    fr_t acc;
    fr_set_zero(&acc);
    for (size_t j = n; j-- > 0;) {
        fr_t cj = c[revbits(j, log)];
        if (regular) fr_to_mont(&cj, &cj);
        fr_mul(&acc, &acc, zz);
        fr_add(&acc, &acc, &cj);
    }
    memcpy(out, acc.l, 32);
}
/* A(z) for the polynomial of degree < n with A(w^i) = a[i] (i < n_in) and 0 (n_in <= i < n), by the barycentric formula
 * A(z) = (z^n - 1) / n * sum_i a[i] w^i / (z - w^i), one batch inversion.  z must not lie in the domain.  Montgomery in and out.  */
void orc_eval_lagrange(const uint64_t *a_in, size_t n_in, size_t n, const uint64_t *z, uint64_t *out) {
    const unsigned log = log2u(n);
    const fr_t *a = (const fr_t *)a_in, *zz = (const fr_t *)z;
    fr_t w, shift;
    domain_roots(log, 0, &w, &shift);
    fr_t *den = (fr_t *)malloc(n_in * sizeof(fr_t)), *pre = (fr_t *)malloc((n_in + 1) * sizeof(fr_t)), *wi = (fr_t *)malloc(n_in * sizeof(fr_t));
    fr_t t;
    fr_set_one(&t);
    fr_set_one(&pre[0]);
    for (size_t i = 0; i < n_in; i++) { /* den[i] = z - w^i; prefix products */
        wi[i] = t;
        fr_sub(&den[i], zz, &t);
        fr_mul(&pre[i + 1], &pre[i], &den[i]);
        fr_mul(&t, &t, &w);
    }
    fr_t inv_all, acc;
    fr_inv(&inv_all, &pre[n_in]);
    fr_set_zero(&acc);
    for (size_t i = n_in; i-- > 0;) {
        fr_t inv_i, term;
        fr_mul(&inv_i, &inv_all, &pre[i]); /* 1 / den[i] */
        fr_mul(&inv_all, &inv_all, &den[i]);
        fr_mul(&term, &a[i], &wi[i]);
        fr_mul(&term, &term, &inv_i);
        fr_add(&acc, &acc, &term);
    }
    fr_t zn, one, nn, n_inv;
    fr_pow2k(&zn, zz, log);
    fr_set_one(&one);
    fr_sub(&zn, &zn, &one);
    fr_set_u64(&nn, (uint64_t)n);
    fr_inv(&n_inv, &nn);
    fr_mul(&acc, &acc, &zn);
    fr_mul(&acc, &acc, &n_inv);
    memcpy(out, acc.l, 32);
    free(den);
    free(pre);
    free(wi);
}
