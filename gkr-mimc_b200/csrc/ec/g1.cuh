// BN254 G1 (y^2 = x^3 + 3 over Fp) for the multi-exponentiation kernels.
//
// Replaces gnark-crypto's ecc/bn254 G1Affine / g1JacExtended arithmetic (reference go.mod:7, un-vendored) behind
// G1Affine.MultiExp, called from prover/gadget/hints.go:182-183 and prover/gadget/prove.go:76,91,189,202,221.
// Bucket sums live in extended Jacobian ("XYZZ") coordinates: x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2, ZZ = 0 <=> infinity -- the
// cheapest known mixed addition (8 products + 2 squarings, no inversion).  Formulas: Explicit-Formulas Database, short
// Weierstrass "xyzz" (Sutherland 2008): madd-2008-s, add-2008-s, dbl-2008-s-1, mdbl-2008-s-1, with a = 0.  Every exceptional
// case (either operand at infinity, equal points, opposite points) is handled: bucket contents come from the caller.
// The affine result of a sum of points is unique, so whatever the coordinates the final bytes equal the reference's.
#pragma once
#include "field.cuh"

namespace ec {

typedef FpMod Fp;

struct G1Affine {  // registers; memory image = gnark-crypto bn254.G1Affine = 8 x u64 (X, Y), Montgomery, infinity = (0, 0)
    Big8 x, y;
};
struct G1X {  // XYZZ
    Big8 x, y, zz, zzz;
};
struct alignas(16) G1XRaw {  // memory image of a G1X: 16 x u64
    uint64_t w[16];
};

EC_HD bool g1_aff_is_inf(const G1Affine& p) { return big_is_zero(p.x) && big_is_zero(p.y); }
EC_HD G1X g1x_inf() {
    G1X r;
    r.x = f_one<Fp>(), r.y = f_one<Fp>(), r.zz = big_zero(), r.zzz = big_zero();
    return r;
}
EC_HD bool g1x_is_inf(const G1X& p) { return big_is_zero(p.zz); }
EC_HD G1Affine g1_aff_load(const uint64_t* p) {
    G1Affine r;
    r.x = big_load(p), r.y = big_load(p + 4);
    return r;
}
EC_HD void g1_aff_store(uint64_t* p, const G1Affine& a) {
    big_store(p, a.x);
    big_store(p + 4, a.y);
}
EC_HD G1X g1x_load(const G1XRaw* p) {
    G1X r;
    r.x = big_load(p->w), r.y = big_load(p->w + 4), r.zz = big_load(p->w + 8), r.zzz = big_load(p->w + 12);
    return r;
}
EC_HD void g1x_store(G1XRaw* p, const G1X& a) {
    big_store(p->w, a.x);
    big_store(p->w + 4, a.y);
    big_store(p->w + 8, a.zz);
    big_store(p->w + 12, a.zzz);
}
EC_HD G1X g1x_from_affine(const G1Affine& p) {
    if (g1_aff_is_inf(p)) return g1x_inf();
    G1X r;
    r.x = p.x, r.y = p.y, r.zz = f_one<Fp>(), r.zzz = f_one<Fp>();
    return r;
}

// The multiplier policy: the bucket-accumulation kernel inlines it (its loop body is ten products; the call ABI of an
// out-of-line multiplier costs ~20 % on the multiplier's own pipe), the short single-thread kernels call one copy.
struct MulInline {
    static EC_HD Big8 mul(const Big8& a, const Big8& b) { return f_mul<Fp>(a, b); }
    static EC_HD Big8 sqr(const Big8& a) { return f_sqr<Fp>(a); }
};
struct MulCall {
    static EC_HD Big8 mul(const Big8& a, const Big8& b) { return f_mulc<Fp>(a, b); }
    static EC_HD Big8 sqr(const Big8& a) { return f_mulc<Fp>(a, a); }
};

// 2 * (affine point), mdbl-2008-s-1 with a = 0.  y = 0 does not occur on this curve (x^3 + 3 = 0 has no point of order 2 in the
// prime-order group G1), but 2*(x, 0) = infinity is returned anyway.
template <class M>
EC_HD G1X g1x_dbl_affine(const G1Affine& p) {
    if (g1_aff_is_inf(p) || big_is_zero(p.y)) return g1x_inf();
    const Big8 u = f_dbl<Fp>(p.y);
    const Big8 v = M::sqr(u);
    const Big8 w = M::mul(u, v);
    const Big8 s = M::mul(p.x, v);
    const Big8 xx = M::sqr(p.x);
    const Big8 m = f_add<Fp>(f_dbl<Fp>(xx), xx);
    G1X r;
    r.x = f_sub<Fp>(f_sub<Fp>(M::sqr(m), s), s);
    r.y = f_sub<Fp>(M::mul(m, f_sub<Fp>(s, r.x)), M::mul(w, p.y));
    r.zz = v;
    r.zzz = w;
    return r;
}
// 2 * P, dbl-2008-s-1 with a = 0
template <class M>
EC_HD G1X g1x_dbl(const G1X& p) {
    if (g1x_is_inf(p) || big_is_zero(p.y)) return g1x_inf();
    const Big8 u = f_dbl<Fp>(p.y);
    const Big8 v = M::sqr(u);
    const Big8 w = M::mul(u, v);
    const Big8 s = M::mul(p.x, v);
    const Big8 xx = M::sqr(p.x);
    const Big8 m = f_add<Fp>(f_dbl<Fp>(xx), xx);
    G1X r;
    r.x = f_sub<Fp>(f_sub<Fp>(M::sqr(m), s), s);
    r.y = f_sub<Fp>(M::mul(m, f_sub<Fp>(s, r.x)), M::mul(w, p.y));
    r.zz = M::mul(v, p.zz);
    r.zzz = M::mul(w, p.zzz);
    return r;
}
// acc + (affine q), madd-2008-s
template <class M>
EC_HD G1X g1x_add_affine(const G1X& acc, const G1Affine& q) {
    if (g1_aff_is_inf(q)) return acc;
    if (g1x_is_inf(acc)) return g1x_from_affine(q);
    const Big8 u2 = M::mul(q.x, acc.zz);
    const Big8 s2 = M::mul(q.y, acc.zzz);
    const Big8 p = f_sub<Fp>(u2, acc.x);
    const Big8 r = f_sub<Fp>(s2, acc.y);
    if (big_is_zero(p)) {
        if (big_is_zero(r)) return g1x_dbl_affine<M>(q);  // same point
        return g1x_inf();                                 // opposite points
    }
    const Big8 pp = M::sqr(p);
    const Big8 ppp = M::mul(p, pp);
    const Big8 qq = M::mul(acc.x, pp);
    G1X o;
    o.x = f_sub<Fp>(f_sub<Fp>(f_sub<Fp>(M::sqr(r), ppp), qq), qq);
    o.y = f_sub<Fp>(M::mul(r, f_sub<Fp>(qq, o.x)), M::mul(acc.y, ppp));
    o.zz = M::mul(acc.zz, pp);
    o.zzz = M::mul(acc.zzz, ppp);
    return o;
}
// a + b, add-2008-s
template <class M>
EC_HD G1X g1x_add(const G1X& a, const G1X& b) {
    if (g1x_is_inf(b)) return a;
    if (g1x_is_inf(a)) return b;
    const Big8 u1 = M::mul(a.x, b.zz);
    const Big8 u2 = M::mul(b.x, a.zz);
    const Big8 s1 = M::mul(a.y, b.zzz);
    const Big8 s2 = M::mul(b.y, a.zzz);
    const Big8 p = f_sub<Fp>(u2, u1);
    const Big8 r = f_sub<Fp>(s2, s1);
    if (big_is_zero(p)) {
        if (big_is_zero(r)) return g1x_dbl<M>(a);
        return g1x_inf();
    }
    const Big8 pp = M::sqr(p);
    const Big8 ppp = M::mul(p, pp);
    const Big8 qq = M::mul(u1, pp);
    G1X o;
    o.x = f_sub<Fp>(f_sub<Fp>(f_sub<Fp>(M::sqr(r), ppp), qq), qq);
    o.y = f_sub<Fp>(M::mul(r, f_sub<Fp>(qq, o.x)), M::mul(s1, ppp));
    o.zz = M::mul(M::mul(a.zz, b.zz), pp);
    o.zzz = M::mul(M::mul(a.zzz, b.zzz), ppp);
    return o;
}
// k * P for a small k (the bucket-chunk offsets of the window reduction), double-and-add from the top bit
template <class M>
EC_HD G1X g1x_mul_small(const G1X& p, uint32_t k) {
    G1X acc = g1x_inf();
    int top = -1;
    for (int i = 0; i < 32; i++)
        if ((k >> i) & 1) top = i;
    for (int i = top; i >= 0; i--) {
        acc = g1x_dbl<M>(acc);
        if ((k >> i) & 1) acc = g1x_add<M>(acc, p);
    }
    return acc;
}
// affine form: x = X/ZZ = X * ZZ^2 / ZZZ^2, y = Y/ZZZ (one inversion)
EC_HD G1Affine g1x_to_affine(const G1X& p) {
    G1Affine r;
    if (g1x_is_inf(p)) {
        r.x = big_zero(), r.y = big_zero();
        return r;
    }
    const Big8 a = f_inv<Fp>(p.zzz);
    const Big8 a2 = f_mulc<Fp>(a, a);
    const Big8 zz2 = f_mulc<Fp>(p.zz, p.zz);
    r.x = f_mulc<Fp>(f_mulc<Fp>(p.x, zz2), a2);
    r.y = f_mulc<Fp>(p.y, a);
    return r;
}

}  // namespace ec
