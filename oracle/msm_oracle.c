/* CPU restatement of the G1 multi-exponentiation / InitialRandomnessHint path -- ORACLE (test infrastructure).
 *
 * Only tests/ and __graft_entry__.smoke() may load this; the product libraries never do.
 *
 * What it follows:
 *   prover/gadget/hints.go:147-159  DeriveRandomnessFromPoint: legacy Keccak-256 of G1Affine.RawBytes(), then fr.SetBytes
 *   prover/gadget/hints.go:162-192  InitialRandomnessHint.Call: two G1Affine.MultiExp + Add
 *   prover/gadget/prove.go:76,91,189,202,221  the other G1Affine.MultiExp call sites of the Groth16 prover
 * The curve arithmetic itself lives in the un-vendored dependency
 *   github.com/consensys/gnark-crypto v0.6.1-0.20220110145513-493bb1c180d9 (reference go.mod:7), ecc/bn254 (G1Affine, MultiExp,
 *   marshal.go RawBytes), ecc/bn254/fp (Element: 4 x u64 little-endian limbs, value * 2^256 mod p, canonical)
 * whose published definitions are restated: BN254 G1 is y^2 = x^3 + 3 over Fp, generator (1, 2), the point at infinity is (0, 0)
 * in affine form, scalars given to MultiExp are fr values in REGULAR form (hints.go:171 FromMont), RawBytes is X || Y big-endian
 * regular form with 0x40 in byte 0 for infinity.  A multi-exponentiation is a group element and its affine form is unique, so
 * parity with the reference is bit-exact whatever algorithm computes it.  THIS oracle computes it the obvious way on purpose --
 * one bit-by-bit double-and-add per point in Jacobian coordinates, summed -- so that it shares nothing with the product's
 * bucket method in XYZZ coordinates.
 *
 * Pinning: the reference holds no golden vector for this path (its tests draw random keys).  The oracle is pinned by published
 * constants instead: Keccak-256 known answers, 2G and q*G = infinity on BN254, and agreement with the independent Python
 * big-integer restatement oracle/pyref_msm.py (tests/test_msm_cpu.py).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fp_t;
typedef struct { fp_t x, y; } aff_t;        /* == gnark-crypto bn254.G1Affine */
typedef struct { fp_t x, y, z; } jac_t;     /* Jacobian, z == 0 <=> infinity */

static const uint64_t FP_P[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t FP_ONE[4] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL};
static const uint64_t FP_R2[4] = {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL};
#define FP_PINV 0x87d20782e4866389ULL /* -p^{-1} mod 2^64 */
static const uint64_t FR_Q[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
#define FR_QINV 0xc2e1f593efffffffULL

static int geq4(const uint64_t t[4], const uint64_t m[4]) {
    for (int i = 3; i >= 0; i--) {
        if (t[i] > m[i]) return 1;
        if (t[i] < m[i]) return 0;
    }
    return 1;
}
static void sub4(uint64_t t[4], const uint64_t m[4]) {
    u128 b = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)t[i] - m[i] - (uint64_t)b;
        t[i] = (uint64_t)d;
        b = (d >> 64) & 1;
    }
}
/* textbook CIOS Montgomery product modulo m (m < 2^254) */
static void mont_mul(uint64_t z[4], const uint64_t x[4], const uint64_t y[4], const uint64_t m[4], uint64_t minv) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)x[j] * y[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t k = t[0] * minv;
        c = (u128)k * m[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)k * m[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || geq4(t, m)) sub4(t, m);
    memcpy(z, t, 32);
}
static void fp_mul(fp_t *z, const fp_t *x, const fp_t *y) { mont_mul(z->l, x->l, y->l, FP_P, FP_PINV); }
static void fp_sqr(fp_t *z, const fp_t *x) { fp_mul(z, x, x); }
static void fp_add(fp_t *z, const fp_t *x, const fp_t *y) {
    uint64_t t[4];
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)x->l[i] + y->l[i];
        t[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq4(t, FP_P)) sub4(t, FP_P);
    memcpy(z->l, t, 32);
}
static void fp_sub(fp_t *z, const fp_t *x, const fp_t *y) {
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
            c += (u128)t[i] + FP_P[i];
            t[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    memcpy(z->l, t, 32);
}
static int fp_is_zero(const fp_t *x) { return (x->l[0] | x->l[1] | x->l[2] | x->l[3]) == 0; }
static int fp_eq(const fp_t *x, const fp_t *y) { return memcmp(x->l, y->l, 32) == 0; }
static void fp_inv(fp_t *z, const fp_t *x) { /* Fermat */
    uint64_t e[4];
    memcpy(e, FP_P, 32);
    e[0] -= 2;
    fp_t acc, base = *x;
    memcpy(acc.l, FP_ONE, 32);
    for (int i = 0; i < 256; i++) {
        if ((e[i / 64] >> (i % 64)) & 1) fp_mul(&acc, &acc, &base);
        fp_sqr(&base, &base);
    }
    *z = acc;
}
static void fp_from_mont(fp_t *z, const fp_t *x) {
    fp_t one = {{1, 0, 0, 0}};
    fp_mul(z, x, &one);
}
static void fp_to_mont(fp_t *z, const fp_t *x) {
    fp_t r2;
    memcpy(r2.l, FP_R2, 32);
    fp_mul(z, x, &r2);
}

/* ---- G1, Jacobian coordinates (x = X/Z^2, y = Y/Z^3), a = 0 */
static int aff_is_inf(const aff_t *p) { return fp_is_zero(&p->x) && fp_is_zero(&p->y); }
static void jac_set_inf(jac_t *p) { memset(p, 0, sizeof *p); }
static void jac_from_aff(jac_t *r, const aff_t *p) {
    if (aff_is_inf(p)) {
        jac_set_inf(r);
        return;
    }
    r->x = p->x;
    r->y = p->y;
    memcpy(r->z.l, FP_ONE, 32);
}
static void jac_dbl(jac_t *r, const jac_t *p) { /* dbl-2009-l */
    if (fp_is_zero(&p->z)) {
        jac_set_inf(r);
        return;
    }
    fp_t a, b, c, d, e, f, t, x3, y3, z3;
    fp_sqr(&a, &p->x);
    fp_sqr(&b, &p->y);
    fp_sqr(&c, &b);
    fp_add(&t, &p->x, &b);
    fp_sqr(&t, &t);
    fp_sub(&t, &t, &a);
    fp_sub(&t, &t, &c);
    fp_add(&d, &t, &t);
    fp_add(&e, &a, &a);
    fp_add(&e, &e, &a);
    fp_sqr(&f, &e);
    fp_sub(&x3, &f, &d);
    fp_sub(&x3, &x3, &d);
    fp_sub(&t, &d, &x3);
    fp_mul(&y3, &e, &t);
    fp_add(&c, &c, &c);
    fp_add(&c, &c, &c);
    fp_add(&c, &c, &c);
    fp_sub(&y3, &y3, &c);
    fp_mul(&z3, &p->y, &p->z);
    fp_add(&z3, &z3, &z3);
    r->x = x3;
    r->y = y3;
    r->z = z3;
}
static void jac_add(jac_t *r, const jac_t *p, const jac_t *q) { /* add-2007-bl with the special cases */
    if (fp_is_zero(&p->z)) {
        *r = *q;
        return;
    }
    if (fp_is_zero(&q->z)) {
        *r = *p;
        return;
    }
    fp_t z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t, x3, y3, z3;
    fp_sqr(&z1z1, &p->z);
    fp_sqr(&z2z2, &q->z);
    fp_mul(&u1, &p->x, &z2z2);
    fp_mul(&u2, &q->x, &z1z1);
    fp_mul(&s1, &p->y, &q->z);
    fp_mul(&s1, &s1, &z2z2);
    fp_mul(&s2, &q->y, &p->z);
    fp_mul(&s2, &s2, &z1z1);
    if (fp_eq(&u1, &u2)) {
        if (fp_eq(&s1, &s2)) jac_dbl(r, p);
        else jac_set_inf(r);
        return;
    }
    fp_sub(&h, &u2, &u1);
    fp_add(&i, &h, &h);
    fp_sqr(&i, &i);
    fp_mul(&j, &h, &i);
    fp_sub(&rr, &s2, &s1);
    fp_add(&rr, &rr, &rr);
    fp_mul(&v, &u1, &i);
    fp_sqr(&x3, &rr);
    fp_sub(&x3, &x3, &j);
    fp_sub(&x3, &x3, &v);
    fp_sub(&x3, &x3, &v);
    fp_sub(&t, &v, &x3);
    fp_mul(&y3, &rr, &t);
    fp_mul(&t, &s1, &j);
    fp_add(&t, &t, &t);
    fp_sub(&y3, &y3, &t);
    fp_add(&z3, &p->z, &q->z);
    fp_sqr(&z3, &z3);
    fp_sub(&z3, &z3, &z1z1);
    fp_sub(&z3, &z3, &z2z2);
    fp_mul(&z3, &z3, &h);
    r->x = x3;
    r->y = y3;
    r->z = z3;
}
static void jac_to_aff(aff_t *r, const jac_t *p) {
    if (fp_is_zero(&p->z)) {
        memset(r, 0, sizeof *r);
        return;
    }
    fp_t zi, zi2, zi3;
    fp_inv(&zi, &p->z);
    fp_sqr(&zi2, &zi);
    fp_mul(&zi3, &zi2, &zi);
    fp_mul(&r->x, &p->x, &zi2);
    fp_mul(&r->y, &p->y, &zi3);
}
/* k * P, k given as 4 x u64 regular-form limbs; bit by bit from the top */
static void jac_scalar_mul(jac_t *r, const aff_t *p, const uint64_t k[4]) {
    jac_t acc, base;
    jac_set_inf(&acc);
    jac_from_aff(&base, p);
    for (int i = 255; i >= 0; i--) {
        jac_dbl(&acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) jac_add(&acc, &acc, &base);
    }
    *r = acc;
}

/* ---- exported API (uint64_t* = Go memory images: G1Affine = 8 words, fr.Element = 4 words) */
int orc_g1_is_on_curve(const uint64_t *pt) {
    aff_t p;
    memcpy(&p, pt, 64);
    if (aff_is_inf(&p)) return 1;
    if (geq4(p.x.l, FP_P) || geq4(p.y.l, FP_P)) return 0;
    fp_t y2, x3, three = {{3, 0, 0, 0}};
    fp_to_mont(&three, &three);
    fp_sqr(&y2, &p.y);
    fp_sqr(&x3, &p.x);
    fp_mul(&x3, &x3, &p.x);
    fp_add(&x3, &x3, &three);
    return fp_eq(&y2, &x3);
}
void orc_g1_generator(uint64_t *out) {
    aff_t g = {{{1, 0, 0, 0}}, {{2, 0, 0, 0}}};
    fp_to_mont(&g.x, &g.x);
    fp_to_mont(&g.y, &g.y);
    memcpy(out, &g, 64);
}
void orc_g1_add(const uint64_t *a, const uint64_t *b, uint64_t *out) { /* G1Affine.Add */
    aff_t pa, pb, r;
    jac_t ja, jb, jr;
    memcpy(&pa, a, 64);
    memcpy(&pb, b, 64);
    jac_from_aff(&ja, &pa);
    jac_from_aff(&jb, &pb);
    jac_add(&jr, &ja, &jb);
    jac_to_aff(&r, &jr);
    memcpy(out, &r, 64);
}
void orc_g1_neg(const uint64_t *a, uint64_t *out) {
    aff_t p;
    memcpy(&p, a, 64);
    if (!aff_is_inf(&p)) {
        fp_t zero = {{0, 0, 0, 0}};
        fp_sub(&p.y, &zero, &p.y);
    }
    memcpy(out, &p, 64);
}
void orc_g1_scalar_mul(const uint64_t *pt, const uint64_t *k_regular, uint64_t *out) { /* G1Affine.ScalarMultiplication */
    aff_t p, r;
    jac_t j;
    memcpy(&p, pt, 64);
    jac_scalar_mul(&j, &p, k_regular);
    jac_to_aff(&r, &j);
    memcpy(out, &r, 64);
}
void orc_fr_from_mont(const uint64_t *x, uint64_t *out) {
    uint64_t one[4] = {1, 0, 0, 0};
    mont_mul(out, x, one, FR_Q, FR_QINV);
}
void orc_fp_from_mont(const uint64_t *x, uint64_t *out) {
    uint64_t one[4] = {1, 0, 0, 0};
    mont_mul(out, x, one, FP_P, FP_PINV);
}
void orc_fp_to_mont(const uint64_t *x, uint64_t *out) { mont_mul(out, x, FP_R2, FP_P, FP_PINV); }

typedef struct {
    const uint64_t *points, *scalars;
    size_t lo, hi;
    int scalars_mont;
    jac_t sum;
} msm_job;
static void *msm_worker(void *arg) {
    msm_job *j = (msm_job *)arg;
    jac_set_inf(&j->sum);
    for (size_t i = j->lo; i < j->hi; i++) {
        aff_t p;
        uint64_t k[4];
        memcpy(&p, j->points + 8 * i, 64);
        if (j->scalars_mont) orc_fr_from_mont(j->scalars + 4 * i, k);
        else memcpy(k, j->scalars + 4 * i, 32);
        jac_t t;
        jac_scalar_mul(&t, &p, k);
        jac_add(&j->sum, &j->sum, &t);
    }
    return NULL;
}
/* G1Affine.MultiExp(points, scalars): sum_i scalars[i] * points[i].  scalars_mont = 0: regular form (what hints.go:171 passes) */
void orc_g1_multiexp(const uint64_t *points, const uint64_t *scalars, size_t n, int scalars_mont, int threads, uint64_t *out) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    msm_job *jobs = (msm_job *)calloc((size_t)threads, sizeof(msm_job));
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++) {
        jobs[t].points = points;
        jobs[t].scalars = scalars;
        jobs[t].scalars_mont = scalars_mont;
        jobs[t].lo = n * (size_t)t / (size_t)threads;
        jobs[t].hi = n * (size_t)(t + 1) / (size_t)threads;
        pthread_create(&th[t], NULL, msm_worker, &jobs[t]);
    }
    jac_t acc;
    jac_set_inf(&acc);
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        jac_add(&acc, &acc, &jobs[t].sum);
    }
    aff_t r;
    jac_to_aff(&r, &acc);
    memcpy(out, &r, 64);
    free(jobs);
    free(th);
}

/* Test bases with a known discrete log: P_i = (a + i*b) * G for i < n (a, b regular-form fr limbs), affine, Montgomery.
 * Then sum_i s_i P_i = (sum_i s_i (a + i b) mod q) * G, which gives a size-independent check of any multi-exponentiation. */
void orc_g1_gen_points(size_t n, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    uint64_t g[8];
    orc_g1_generator(g);
    aff_t ga, step_a;
    memcpy(&ga, g, 64);
    jac_t cur, step;
    jac_scalar_mul(&cur, &ga, a);
    jac_scalar_mul(&step, &ga, b);
    jac_to_aff(&step_a, &step);
    jac_from_aff(&step, &step_a);
    enum { BLK = 512 };
    jac_t *blk = (jac_t *)malloc(sizeof(jac_t) * BLK);
    fp_t *pref = (fp_t *)malloc(sizeof(fp_t) * BLK);
    for (size_t base = 0; base < n; base += BLK) {
        size_t m = n - base < BLK ? n - base : BLK;
        for (size_t i = 0; i < m; i++) {
            blk[i] = cur;
            jac_add(&cur, &cur, &step);
        }
        /* batch inversion of the non-zero z (Montgomery's trick) */
        fp_t acc;
        memcpy(acc.l, FP_ONE, 32);
        for (size_t i = 0; i < m; i++) {
            pref[i] = acc;
            if (!fp_is_zero(&blk[i].z)) fp_mul(&acc, &acc, &blk[i].z);
        }
        fp_t inv;
        fp_inv(&inv, &acc);
        for (size_t i = m; i-- > 0;) {
            aff_t r;
            if (fp_is_zero(&blk[i].z)) {
                memset(&r, 0, sizeof r);
            } else {
                fp_t zi, zi2, zi3;
                fp_mul(&zi, &inv, &pref[i]);
                fp_mul(&inv, &inv, &blk[i].z);
                fp_sqr(&zi2, &zi);
                fp_mul(&zi3, &zi2, &zi);
                fp_mul(&r.x, &blk[i].x, &zi2);
                fp_mul(&r.y, &blk[i].y, &zi3);
            }
            memcpy(out + 8 * (base + i), &r, 64);
        }
    }
    free(blk);
    free(pref);
}

/* ---- legacy Keccak-256 (golang.org/x/crypto/sha3 NewLegacyKeccak256: rate 136, domain byte 0x01) */
static const uint64_t KRC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
                                 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
                                 0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
                                 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
                                 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14}; /* [x + 5y] */
static uint64_t rol64(uint64_t v, int n) { return n ? (v << n) | (v >> (64 - n)) : v; }
static void keccak_f(uint64_t s[25]) { /* lane (x, y) at s[x + 5y] */
    for (int rnd = 0; rnd < 24; rnd++) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) s[i] ^= d[i % 5];
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(s[x + 5 * y], KROT[x + 5 * y]);
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) s[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        s[0] ^= KRC[rnd];
    }
}
void orc_keccak256(const uint8_t *data, size_t len, uint8_t out[32]) {
    uint64_t s[25];
    memset(s, 0, sizeof s);
    uint8_t blk[136];
    size_t off = 0;
    for (;;) {
        size_t m = len - off < 136 ? len - off : 136;
        int last = m < 136;
        memset(blk, 0, 136);
        memcpy(blk, data + off, m);
        if (last) {
            blk[m] ^= 0x01;
            blk[135] ^= 0x80;
        }
        for (int i = 0; i < 17; i++) {
            uint64_t w = 0;
            for (int k = 7; k >= 0; k--) w = (w << 8) | blk[8 * i + k];
            s[i] ^= w;
        }
        keccak_f(s);
        off += m;
        if (last) break;
    }
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(s[i] >> (8 * k));
}
/* G1Affine.RawBytes */
void orc_g1_raw_bytes(const uint64_t *pt, uint8_t out[64]) {
    aff_t p;
    memcpy(&p, pt, 64);
    memset(out, 0, 64);
    if (aff_is_inf(&p)) {
        out[0] = 0x40;
        return;
    }
    fp_t x, y;
    fp_from_mont(&x, &p.x);
    fp_from_mont(&y, &p.y);
    for (int i = 0; i < 32; i++) {
        out[31 - i] = (uint8_t)(x.l[i / 8] >> (8 * (i % 8)));
        out[63 - i] = (uint8_t)(y.l[i / 8] >> (8 * (i % 8)));
    }
}
/* hints.go:147-159 -> fr value in REGULAR form (what initialRandomness.ToBigIntRegular(oups[0]) hands the solver) */
void orc_derive_randomness_from_point(const uint64_t *pt, uint64_t out[4]) {
    uint8_t raw[64], h[32];
    orc_g1_raw_bytes(pt, raw);
    orc_keccak256(raw, 64, h);
    uint64_t v[4];
    for (int i = 0; i < 4; i++) {
        v[i] = 0;
        for (int k = 0; k < 8; k++) v[i] |= (uint64_t)h[31 - (8 * i + k)] << (8 * k);
    }
    while (geq4(v, FR_Q)) sub4(v, FR_Q); /* fr.SetBytes: big-endian integer mod q (2^256 / q < 6) */
    memcpy(out, v, 32);
}
/* hints.go:162-192 */
void orc_initial_randomness(const uint64_t *pub_points, const uint64_t *pub_scalars, size_t n_pub, const uint64_t *priv_points,
                            const uint64_t *priv_scalars, size_t n_priv, int scalars_mont, int threads, uint64_t *krs_gkr_priv, uint64_t *randomness) {
    uint64_t krs[8];
    orc_g1_multiexp(pub_points, pub_scalars, n_pub, scalars_mont, threads, krs);
    orc_g1_multiexp(priv_points, priv_scalars, n_priv, scalars_mont, threads, krs_gkr_priv);
    orc_g1_add(krs, krs_gkr_priv, krs);
    orc_derive_randomness_from_point(krs, randomness);
}

/* ---- G1Affine.MultiExp as gnark-crypto computes it (ecc/bn254/multiexp.go, published algorithm restated): the bucket method.
 * Used as the CPU BASELINE of the multi-exponentiation (tools/gpu_groth16_side.py) and as a fast second oracle at sizes where one
 * double-and-add per point takes minutes; tests/test_msm_cpu.py pins it to orc_g1_multiexp above.
 *   bestC:            c = argmin over the implemented widths {4, 5, 8, 16} of 256 * (n + 2^c) / c
 *   partitionScalars: signed c-bit digits: a digit >= 2^(c-1) becomes digit - 2^c with a carry into the next window
 *   one task per window (a goroutine per chunk there, a pthread here): bucket[|d| - 1] += / -= point (mixed addition), then
 *                     the running-sum reduction  sum_b (b + 1) bucket[b]  from the top bucket down
 *   msmReduceChunk:   Horner over the windows from the top: c doublings, then add the window's sum
 * Coordinates here are plain Jacobian with the generic addition (the reference uses extended Jacobian buckets; the product XYZZ):
 * the group elements are the same.                                                                                                  */
typedef struct {
    const uint64_t *points;
    const int32_t *digits; /* [n][W] */
    size_t n;
    unsigned W, c;
    volatile int *next_window;
    jac_t *window_sum; /* [W] */
} bkt_job;
static void *bkt_worker(void *arg) {
    bkt_job *j = (bkt_job *)arg;
    const size_t nb = (size_t)1 << (j->c - 1);
    jac_t *bucket = (jac_t *)malloc(sizeof(jac_t) * nb);
    for (;;) {
        const int w = __sync_fetch_and_add(j->next_window, 1);
        if (w >= (int)j->W) break;
        for (size_t b = 0; b < nb; b++) jac_set_inf(&bucket[b]);
        for (size_t i = 0; i < j->n; i++) {
            const int32_t d = j->digits[i * j->W + (size_t)w];
            if (d == 0) continue;
            aff_t p;
            memcpy(&p, j->points + 8 * i, 64);
            if (aff_is_inf(&p)) continue;
            if (d < 0) {
                fp_t zero = {{0, 0, 0, 0}};
                fp_sub(&p.y, &zero, &p.y);
            }
            jac_t pj;
            jac_from_aff(&pj, &p);
            const size_t b = (size_t)(d < 0 ? -d : d) - 1;
            jac_add(&bucket[b], &bucket[b], &pj);
        }
        jac_t run, total;
        jac_set_inf(&run);
        jac_set_inf(&total);
        for (size_t b = nb; b-- > 0;) {
            jac_add(&run, &run, &bucket[b]);
            jac_add(&total, &total, &run);
        }
        j->window_sum[w] = total;
    }
    free(bucket);
    return NULL;
}
void orc_g1_multiexp_buckets(const uint64_t *points, const uint64_t *scalars, size_t n, int scalars_mont, int threads, uint64_t *out) {
    if (n == 0) {
        memset(out, 0, 64);
        return;
    }
    static const unsigned implemented[4] = {4, 5, 8, 16};
    unsigned c = 4;
    double best = 1e300;
    for (int k = 0; k < 4; k++) {
        const double cost = 256.0 * ((double)n + (double)((size_t)1 << implemented[k])) / (double)implemented[k];
        if (cost < best) best = cost, c = implemented[k];
    }
    const unsigned W = 254 / c + 1; /* W * c >= 255: the top digit never carries out for a scalar below q < 2^254 */
    int32_t *digits = (int32_t *)malloc(sizeof(int32_t) * n * W);
    for (size_t i = 0; i < n; i++) {
        uint64_t k[5] = {0, 0, 0, 0, 0};
        if (scalars_mont) orc_fr_from_mont(scalars + 4 * i, k);
        else memcpy(k, scalars + 4 * i, 32);
        uint32_t carry = 0;
        for (unsigned w = 0; w < W; w++) {
            const unsigned pos = w * c, word = pos >> 6, off = pos & 63;
            uint64_t raw = 0;
            if (word < 4) {
                raw = k[word] >> off;
                if (off + c > 64) raw |= k[word + 1] << (64 - off);
                raw &= ((uint64_t)1 << c) - 1;
            }
            raw += carry;
            if (raw >= ((uint64_t)1 << (c - 1))) {
                digits[i * W + w] = (int32_t)((int64_t)raw - ((int64_t)1 << c));
                carry = 1;
            } else {
                digits[i * W + w] = (int32_t)raw;
                carry = 0;
            }
        }
    }
    if (threads < 1) threads = 1;
    if ((unsigned)threads > W) threads = (int)W;
    jac_t *wsum = (jac_t *)malloc(sizeof(jac_t) * W);
    volatile int next = 0;
    bkt_job job = {points, digits, n, W, c, &next, wsum};
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, bkt_worker, &job);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    jac_t acc;
    jac_set_inf(&acc);
    for (unsigned w = W; w-- > 0;) {
        for (unsigned d = 0; d < c; d++) jac_dbl(&acc, &acc);
        jac_add(&acc, &acc, &wsum[w]);
    }
    aff_t r;
    jac_to_aff(&r, &acc);
    memcpy(out, &r, 64);
    free(digits);
    free(wsum);
    free(th);
}

/* =================================================================================================================================
 * G2: y^2 = x^3 + 3/(9+u) over Fp2 = Fp[u]/(u^2+1) -- G2Affine.MultiExp (prover/gadget/prove.go:277, Bs).
 * gnark-crypto bn254.G2Affine = {X, Y fptower.E2}, E2 = {A0, A1 fp.Element}: 16 words, Montgomery, infinity all zero.
 * Fp2 products are SCHOOLBOOK here (four base products; the product library uses Karatsuba and complex squaring) and the group law
 * is Jacobian (the product uses XYZZ), one double-and-add per point, summed.
 * Pinning: the published generator of G2 (EIP-197 / gnark-crypto) must satisfy the curve equation and have order q
 * (tests/test_msm_cpu.py), and agreement with the Python big-integer restatement in oracle/pyref_msm.py.
 * ================================================================================================================================= */
typedef struct { fp_t a0, a1; } f2_t;
typedef struct { f2_t x, y; } aff2_t;
typedef struct { f2_t x, y, z; } jac2_t;

static void f2_add(f2_t *z, const f2_t *x, const f2_t *y) { fp_add(&z->a0, &x->a0, &y->a0); fp_add(&z->a1, &x->a1, &y->a1); }
static void f2_sub(f2_t *z, const f2_t *x, const f2_t *y) { fp_sub(&z->a0, &x->a0, &y->a0); fp_sub(&z->a1, &x->a1, &y->a1); }
static void f2_mul(f2_t *z, const f2_t *x, const f2_t *y) {
    fp_t t0, t1, t2, t3;
    fp_mul(&t0, &x->a0, &y->a0);
    fp_mul(&t1, &x->a1, &y->a1);
    fp_mul(&t2, &x->a0, &y->a1);
    fp_mul(&t3, &x->a1, &y->a0);
    fp_sub(&z->a0, &t0, &t1);
    fp_add(&z->a1, &t2, &t3);
}
static void f2_sqr(f2_t *z, const f2_t *x) { f2_t t = *x; f2_mul(z, &t, &t); }
static int f2_is_zero(const f2_t *x) { return fp_is_zero(&x->a0) && fp_is_zero(&x->a1); }
static int f2_eq(const f2_t *x, const f2_t *y) { return fp_eq(&x->a0, &y->a0) && fp_eq(&x->a1, &y->a1); }
static void f2_inv(f2_t *z, const f2_t *x) { /* conj(x) / (a0^2 + a1^2) */
    fp_t n, t, ni, zero = {{0, 0, 0, 0}};
    fp_sqr(&n, &x->a0);
    fp_sqr(&t, &x->a1);
    fp_add(&n, &n, &t);
    fp_inv(&ni, &n);
    fp_mul(&z->a0, &x->a0, &ni);
    fp_mul(&t, &x->a1, &ni);
    fp_sub(&z->a1, &zero, &t);
}
static void f2_set_one(f2_t *z) { memcpy(z->a0.l, FP_ONE, 32); memset(z->a1.l, 0, 32); }

static int aff2_is_inf(const aff2_t *p) { return f2_is_zero(&p->x) && f2_is_zero(&p->y); }
static void jac2_set_inf(jac2_t *p) { memset(p, 0, sizeof *p); }
static void jac2_from_aff(jac2_t *r, const aff2_t *p) {
    if (aff2_is_inf(p)) { jac2_set_inf(r); return; }
    r->x = p->x;
    r->y = p->y;
    f2_set_one(&r->z);
}
static void jac2_dbl(jac2_t *r, const jac2_t *p) { /* dbl-2009-l, a = 0 */
    if (f2_is_zero(&p->z)) { jac2_set_inf(r); return; }
    f2_t a, b, c, d, e, f, t, x3, y3, z3;
    f2_sqr(&a, &p->x);
    f2_sqr(&b, &p->y);
    f2_sqr(&c, &b);
    f2_add(&t, &p->x, &b);
    f2_sqr(&t, &t);
    f2_sub(&t, &t, &a);
    f2_sub(&t, &t, &c);
    f2_add(&d, &t, &t);
    f2_add(&e, &a, &a);
    f2_add(&e, &e, &a);
    f2_sqr(&f, &e);
    f2_sub(&x3, &f, &d);
    f2_sub(&x3, &x3, &d);
    f2_sub(&t, &d, &x3);
    f2_mul(&y3, &e, &t);
    f2_add(&c, &c, &c);
    f2_add(&c, &c, &c);
    f2_add(&c, &c, &c);
    f2_sub(&y3, &y3, &c);
    f2_mul(&z3, &p->y, &p->z);
    f2_add(&z3, &z3, &z3);
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac2_add(jac2_t *r, const jac2_t *p, const jac2_t *q) { /* add-2007-bl with the special cases */
    if (f2_is_zero(&p->z)) { *r = *q; return; }
    if (f2_is_zero(&q->z)) { *r = *p; return; }
    f2_t z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t, x3, y3, z3;
    f2_sqr(&z1z1, &p->z);
    f2_sqr(&z2z2, &q->z);
    f2_mul(&u1, &p->x, &z2z2);
    f2_mul(&u2, &q->x, &z1z1);
    f2_mul(&s1, &p->y, &q->z);
    f2_mul(&s1, &s1, &z2z2);
    f2_mul(&s2, &q->y, &p->z);
    f2_mul(&s2, &s2, &z1z1);
    if (f2_eq(&u1, &u2)) {
        if (f2_eq(&s1, &s2)) jac2_dbl(r, p);
        else jac2_set_inf(r);
        return;
    }
    f2_sub(&h, &u2, &u1);
    f2_add(&i, &h, &h);
    f2_sqr(&i, &i);
    f2_mul(&j, &h, &i);
    f2_sub(&rr, &s2, &s1);
    f2_add(&rr, &rr, &rr);
    f2_mul(&v, &u1, &i);
    f2_sqr(&x3, &rr);
    f2_sub(&x3, &x3, &j);
    f2_sub(&x3, &x3, &v);
    f2_sub(&x3, &x3, &v);
    f2_sub(&t, &v, &x3);
    f2_mul(&y3, &rr, &t);
    f2_mul(&t, &s1, &j);
    f2_add(&t, &t, &t);
    f2_sub(&y3, &y3, &t);
    f2_add(&z3, &p->z, &q->z);
    f2_sqr(&z3, &z3);
    f2_sub(&z3, &z3, &z1z1);
    f2_sub(&z3, &z3, &z2z2);
    f2_mul(&z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac2_to_aff(aff2_t *r, const jac2_t *p) {
    if (f2_is_zero(&p->z)) { memset(r, 0, sizeof *r); return; }
    f2_t zi, zi2, zi3;
    f2_inv(&zi, &p->z);
    f2_sqr(&zi2, &zi);
    f2_mul(&zi3, &zi2, &zi);
    f2_mul(&r->x, &p->x, &zi2);
    f2_mul(&r->y, &p->y, &zi3);
}
static void jac2_scalar_mul(jac2_t *r, const aff2_t *p, const uint64_t k[4]) {
    jac2_t acc, base;
    jac2_set_inf(&acc);
    jac2_from_aff(&base, p);
    for (int i = 255; i >= 0; i--) {
        jac2_dbl(&acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) jac2_add(&acc, &acc, &base);
    }
    *r = acc;
}

/* the generator of G2 (EIP-197 / gnark-crypto bn254.Generators), regular form, little-endian limbs: X.A0, X.A1, Y.A0, Y.A1 */
static const uint64_t G2_GEN[16] = {
    0x46debd5cd992f6edULL, 0x674322d4f75edaddULL, 0x426a00665e5c4479ULL, 0x1800deef121f1e76ULL, /* 10857046999023057135944570762232829481370756359578518086990519993285655852781 */
    0x97e485b7aef312c2ULL, 0xf1aa493335a9e712ULL, 0x7260bfb731fb5d25ULL, 0x198e9393920d483aULL, /* 11559732032986387107991004021392285783925812861821192530917403151452391805634 */
    0x4ce6cc0166fa7daaULL, 0xe3d1e7690c43d37bULL, 0x4aab71808dcb408fULL, 0x12c85ea5db8c6debULL, /* 8495653923123431417604973247489272438418190587263600148770280649306958101930 */
    0x55acdadcd122975bULL, 0xbc4b313370b38ef3ULL, 0xec9e99ad690c3395ULL, 0x090689d0585ff075ULL, /* 4082367875863433681332203403145435568316851327593401208105741076214120093531 */
};
void orc_g2_generator(uint64_t *out) {
    for (int k = 0; k < 4; k++) orc_fp_to_mont(G2_GEN + 4 * k, out + 4 * k);
}
int orc_g2_is_on_curve(const uint64_t *pt) { /* y^2 = x^3 + 3/(9+u) */
    aff2_t p;
    memcpy(&p, pt, 128);
    if (aff2_is_inf(&p)) return 1;
    f2_t nine_u, three, b, y2, x3;
    fp_t nine = {{9, 0, 0, 0}}, one = {{1, 0, 0, 0}}, thr = {{3, 0, 0, 0}};
    fp_to_mont(&nine_u.a0, &nine);
    fp_to_mont(&nine_u.a1, &one);
    fp_to_mont(&three.a0, &thr);
    memset(three.a1.l, 0, 32);
    f2_inv(&b, &nine_u);
    f2_mul(&b, &b, &three);
    f2_sqr(&y2, &p.y);
    f2_sqr(&x3, &p.x);
    f2_mul(&x3, &x3, &p.x);
    f2_add(&x3, &x3, &b);
    return f2_eq(&y2, &x3);
}
void orc_g2_add(const uint64_t *a, const uint64_t *b, uint64_t *out) {
    aff2_t pa, pb, r;
    jac2_t ja, jb, js;
    memcpy(&pa, a, 128);
    memcpy(&pb, b, 128);
    jac2_from_aff(&ja, &pa);
    jac2_from_aff(&jb, &pb);
    jac2_add(&js, &ja, &jb);
    jac2_to_aff(&r, &js);
    memcpy(out, &r, 128);
}
void orc_g2_neg(const uint64_t *a, uint64_t *out) {
    aff2_t p;
    f2_t zero;
    memcpy(&p, a, 128);
    memset(&zero, 0, sizeof zero);
    if (!aff2_is_inf(&p)) f2_sub(&p.y, &zero, &p.y);
    memcpy(out, &p, 128);
}
void orc_g2_scalar_mul(const uint64_t *pt, const uint64_t *k_regular, uint64_t *out) {
    aff2_t p, r;
    jac2_t j;
    memcpy(&p, pt, 128);
    jac2_scalar_mul(&j, &p, k_regular);
    jac2_to_aff(&r, &j);
    memcpy(out, &r, 128);
}
typedef struct {
    const uint64_t *points, *scalars;
    size_t lo, hi;
    int scalars_mont;
    jac2_t sum;
} msm2_job;
static void *msm2_worker(void *arg) {
    msm2_job *j = (msm2_job *)arg;
    jac2_set_inf(&j->sum);
    for (size_t i = j->lo; i < j->hi; i++) {
        aff2_t p;
        uint64_t k[4];
        memcpy(&p, j->points + 16 * i, 128);
        if (j->scalars_mont) orc_fr_from_mont(j->scalars + 4 * i, k);
        else memcpy(k, j->scalars + 4 * i, 32);
        jac2_t t;
        jac2_scalar_mul(&t, &p, k);
        jac2_add(&j->sum, &j->sum, &t);
    }
    return NULL;
}
/* G2Affine.MultiExp(points, scalars) */
void orc_g2_multiexp(const uint64_t *points, const uint64_t *scalars, size_t n, int scalars_mont, int threads, uint64_t *out) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    msm2_job *jobs = (msm2_job *)calloc((size_t)threads, sizeof(msm2_job));
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++) {
        jobs[t].points = points;
        jobs[t].scalars = scalars;
        jobs[t].scalars_mont = scalars_mont;
        jobs[t].lo = n * (size_t)t / (size_t)threads;
        jobs[t].hi = n * (size_t)(t + 1) / (size_t)threads;
        pthread_create(&th[t], NULL, msm2_worker, &jobs[t]);
    }
    jac2_t acc;
    jac2_set_inf(&acc);
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        jac2_add(&acc, &acc, &jobs[t].sum);
    }
    aff2_t r;
    jac2_to_aff(&r, &acc);
    memcpy(out, &r, 128);
    free(jobs);
    free(th);
}
/* P_i = (a + i*b) * G2gen, i < n: bases with a known discrete log (closed-form check of a multi-exponentiation at any size) */
void orc_g2_gen_points(size_t n, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    uint64_t g[16];
    orc_g2_generator(g);
    aff2_t ga;
    memcpy(&ga, g, 128);
    jac2_t cur, step;
    jac2_scalar_mul(&cur, &ga, a);
    jac2_scalar_mul(&step, &ga, b);
    enum { BLK2 = 256 };
    jac2_t *blk = (jac2_t *)malloc(sizeof(jac2_t) * BLK2);
    f2_t *pref = (f2_t *)malloc(sizeof(f2_t) * BLK2);
    for (size_t base = 0; base < n; base += BLK2) {
        size_t m = n - base < BLK2 ? n - base : BLK2;
        for (size_t i = 0; i < m; i++) {
            blk[i] = cur;
            jac2_add(&cur, &cur, &step);
        }
        f2_t acc;
        f2_set_one(&acc);
        for (size_t i = 0; i < m; i++) {
            pref[i] = acc;
            if (!f2_is_zero(&blk[i].z)) f2_mul(&acc, &acc, &blk[i].z);
        }
        f2_t inv;
        f2_inv(&inv, &acc);
        for (size_t i = m; i-- > 0;) {
            aff2_t r;
            if (f2_is_zero(&blk[i].z)) {
                memset(&r, 0, sizeof r);
            } else {
                f2_t zi, zi2, zi3;
                f2_mul(&zi, &inv, &pref[i]);
                f2_mul(&inv, &inv, &blk[i].z);
                f2_sqr(&zi2, &zi);
                f2_mul(&zi3, &zi2, &zi);
                f2_mul(&r.x, &blk[i].x, &zi2);
                f2_mul(&r.y, &blk[i].y, &zi3);
            }
            memcpy(out + 16 * (base + i), &r, 128);
        }
    }
    free(blk);
    free(pref);
}
