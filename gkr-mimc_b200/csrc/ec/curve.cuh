// BN254 G1 (y^2 = x^3 + 3 over Fp) and G2 (y^2 = x^3 + 3/(9+u) over Fp2 = Fp[u]/(u^2+1)) for the multi-exponentiation kernels:
// one group law, written once over the coordinate field.
//
// Replaces gnark-crypto's ecc/bn254 G1Affine / g1JacExtended and G2Affine / g2JacExtended arithmetic (reference go.mod:7,
// un-vendored) behind G1Affine.MultiExp (prover/gadget/hints.go:182-183, prover/gadget/prove.go:76,91,189,202,221) and
// G2Affine.MultiExp (prove.go:277).
// Bucket sums live in extended Jacobian ("XYZZ") coordinates: x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2, ZZ = 0 <=> infinity -- the
// cheapest known mixed addition (8 products + 2 squarings, no inversion).  Formulas: Explicit-Formulas Database, short
// Weierstrass "xyzz" (Sutherland 2008): madd-2008-s, add-2008-s, dbl-2008-s-1, mdbl-2008-s-1, with a = 0 (both curves; the
// constant b does not appear in any of them).  Every exceptional case (either operand at infinity, equal points, opposite
// points) is handled: bucket contents come from the caller.
// The affine result of a sum of points is unique, so whatever the coordinates the final bytes equal the reference's.
#pragma once
#include "field.cuh"

namespace ec {

typedef FpMod Fp;

// ---- coordinate fields ----------------------------------------------------------------------------------------------------------
// Fp: memory image = fp.Element, 4 x u64
struct FpBase {
    typedef Big8 El;
    static constexpr int WORDS = 4;
    static EC_HD El zero() { return big_zero(); }
    static EC_HD El one() { return f_one<Fp>(); }
    static EC_HD bool is_zero(const El& a) { return big_is_zero(a); }
    static EC_HD El add(const El& a, const El& b) { return f_add<Fp>(a, b); }
    static EC_HD El sub(const El& a, const El& b) { return f_sub<Fp>(a, b); }
    static EC_HD El dbl(const El& a) { return f_dbl<Fp>(a); }
    static EC_HD El neg(const El& a) { return f_neg<Fp>(a); }
    static EC_HD El load(const uint64_t* p) { return big_load(p); }
    static EC_HD void store(uint64_t* p, const El& a) { big_store(p, a); }
    static EC_HD El inv(const El& a) { return f_inv<Fp>(a); }  // one call per multi-exponentiation
    static EC_HD El from_mont(const El& a) { return f_from_mont<Fp>(a); }
};
// Fp2: memory image = gnark-crypto's fptower.E2 {A0, A1 fp.Element}, 8 x u64; value A0 + A1 u, u^2 = -1
struct Fp2El {
    Big8 c0, c1;
};
struct Fp2Base {
    typedef Fp2El El;
    static constexpr int WORDS = 8;
    static EC_HD El zero() {
        El r;
        r.c0 = big_zero(), r.c1 = big_zero();
        return r;
    }
    static EC_HD El one() {
        El r;
        r.c0 = f_one<Fp>(), r.c1 = big_zero();
        return r;
    }
    static EC_HD bool is_zero(const El& a) { return big_is_zero(a.c0) && big_is_zero(a.c1); }
    static EC_HD El add(const El& a, const El& b) {
        El r;
        r.c0 = f_add<Fp>(a.c0, b.c0), r.c1 = f_add<Fp>(a.c1, b.c1);
        return r;
    }
    static EC_HD El sub(const El& a, const El& b) {
        El r;
        r.c0 = f_sub<Fp>(a.c0, b.c0), r.c1 = f_sub<Fp>(a.c1, b.c1);
        return r;
    }
    static EC_HD El dbl(const El& a) { return add(a, a); }
    static EC_HD El neg(const El& a) { return sub(zero(), a); }
    static EC_HD El load(const uint64_t* p) {
        El r;
        r.c0 = big_load(p), r.c1 = big_load(p + 4);
        return r;
    }
    static EC_HD void store(uint64_t* p, const El& a) {
        big_store(p, a.c0);
        big_store(p + 4, a.c1);
    }
    // 1 / (a0 + a1 u) = (a0 - a1 u) / (a0^2 + a1^2)
    static EC_HD El inv(const El& a) {
        const Big8 n = f_add<Fp>(f_mulc<Fp>(a.c0, a.c0), f_mulc<Fp>(a.c1, a.c1));
        const Big8 ni = f_inv<Fp>(n);
        El r;
        r.c0 = f_mulc<Fp>(a.c0, ni), r.c1 = f_neg<Fp>(f_mulc<Fp>(a.c1, ni));
        return r;
    }
    static EC_HD El from_mont(const El& a) {
        El r;
        r.c0 = f_from_mont<Fp>(a.c0), r.c1 = f_from_mont<Fp>(a.c1);
        return r;
    }
};

// The multiplier policy: the bucket-accumulation kernel inlines it (its loop body is ten products; the call ABI of an
// out-of-line multiplier costs ~20 % on the multiplier's own pipe), the short single-thread kernels call one copy.
struct FpMulInline {
    static EC_HD Big8 mul(const Big8& a, const Big8& b) { return f_mul<Fp>(a, b); }
    static EC_HD Big8 sqr(const Big8& a) { return f_sqr<Fp>(a); }
};
struct FpMulCall {
    static EC_HD Big8 mul(const Big8& a, const Big8& b) { return f_mulc<Fp>(a, b); }
    static EC_HD Big8 sqr(const Big8& a) { return f_mulc<Fp>(a, a); }
};
// Fp2 products over an Fp policy: Karatsuba (3 base products), complex squaring (2)
template <class B>
struct Fp2Mul {
    static EC_HD Fp2El mul(const Fp2El& a, const Fp2El& b) {
        const Big8 v0 = B::mul(a.c0, b.c0), v1 = B::mul(a.c1, b.c1);
        const Big8 m = B::mul(f_add<Fp>(a.c0, a.c1), f_add<Fp>(b.c0, b.c1));
        Fp2El r;
        r.c0 = f_sub<Fp>(v0, v1);
        r.c1 = f_sub<Fp>(f_sub<Fp>(m, v0), v1);
        return r;
    }
    static EC_HD Fp2El sqr(const Fp2El& a) {
        const Big8 t = B::mul(a.c0, a.c1);
        Fp2El r;
        r.c0 = B::mul(f_add<Fp>(a.c0, a.c1), f_sub<Fp>(a.c0, a.c1));
        r.c1 = f_dbl<Fp>(t);
        return r;
    }
};

// ---- the group law ------------------------------------------------------------------------------------------------------------
template <class F, class MInl, class MCal>
struct Curve {
    typedef F Field;
    typedef typename F::El El;
    typedef MInl MInline;
    typedef MCal MCall;
    static constexpr int AFF_WORDS = 2 * F::WORDS;  // memory image of an affine point: X then Y; infinity = all zero
    static constexpr int X_WORDS = 4 * F::WORDS;    // memory image of an XYZZ point: X, Y, ZZ, ZZZ
    struct Affine {
        El x, y;
    };
    struct X {  // XYZZ
        El x, y, zz, zzz;
    };

    static EC_HD bool aff_is_inf(const Affine& p) { return F::is_zero(p.x) && F::is_zero(p.y); }
    static EC_HD X x_inf() {
        X r;
        r.x = F::one(), r.y = F::one(), r.zz = F::zero(), r.zzz = F::zero();
        return r;
    }
    static EC_HD bool x_is_inf(const X& p) { return F::is_zero(p.zz); }
    static EC_HD Affine aff_load(const uint64_t* p) {
        Affine r;
        r.x = F::load(p), r.y = F::load(p + F::WORDS);
        return r;
    }
    static EC_HD void aff_store(uint64_t* p, const Affine& a) {
        F::store(p, a.x);
        F::store(p + F::WORDS, a.y);
    }
    static EC_HD Affine aff_neg(const Affine& p) {
        Affine r = p;
        if (!aff_is_inf(p)) r.y = F::neg(p.y);
        return r;
    }
    static EC_HD X x_load(const uint64_t* p) {
        X r;
        r.x = F::load(p), r.y = F::load(p + F::WORDS), r.zz = F::load(p + 2 * F::WORDS), r.zzz = F::load(p + 3 * F::WORDS);
        return r;
    }
    static EC_HD void x_store(uint64_t* p, const X& a) {
        F::store(p, a.x);
        F::store(p + F::WORDS, a.y);
        F::store(p + 2 * F::WORDS, a.zz);
        F::store(p + 3 * F::WORDS, a.zzz);
    }
    static EC_HD X from_affine(const Affine& p) {
        if (aff_is_inf(p)) return x_inf();
        X r;
        r.x = p.x, r.y = p.y, r.zz = F::one(), r.zzz = F::one();
        return r;
    }

    // 2 * (affine point), mdbl-2008-s-1 with a = 0.  y = 0 does not occur in the prime-order groups G1 and G2 (a point with
    // y = 0 has order 2), but 2 * (x, 0) = infinity is returned anyway.
    template <class M>
    static EC_HD X dbl_affine(const Affine& p) {
        if (aff_is_inf(p) || F::is_zero(p.y)) return x_inf();
        const El u = F::dbl(p.y);
        const El v = M::sqr(u);
        const El w = M::mul(u, v);
        const El s = M::mul(p.x, v);
        const El xx = M::sqr(p.x);
        const El m = F::add(F::dbl(xx), xx);
        X r;
        r.x = F::sub(F::sub(M::sqr(m), s), s);
        r.y = F::sub(M::mul(m, F::sub(s, r.x)), M::mul(w, p.y));
        r.zz = v;
        r.zzz = w;
        return r;
    }
    // 2 * P, dbl-2008-s-1 with a = 0
    template <class M>
    static EC_HD X dbl(const X& p) {
        if (x_is_inf(p) || F::is_zero(p.y)) return x_inf();
        const El u = F::dbl(p.y);
        const El v = M::sqr(u);
        const El w = M::mul(u, v);
        const El s = M::mul(p.x, v);
        const El xx = M::sqr(p.x);
        const El m = F::add(F::dbl(xx), xx);
        X r;
        r.x = F::sub(F::sub(M::sqr(m), s), s);
        r.y = F::sub(M::mul(m, F::sub(s, r.x)), M::mul(w, p.y));
        r.zz = M::mul(v, p.zz);
        r.zzz = M::mul(w, p.zzz);
        return r;
    }
    // acc + (affine q), madd-2008-s
    template <class M>
    static EC_HD X add_affine(const X& acc, const Affine& q) {
        if (aff_is_inf(q)) return acc;
        if (x_is_inf(acc)) return from_affine(q);
        const El u2 = M::mul(q.x, acc.zz);
        const El s2 = M::mul(q.y, acc.zzz);
        const El p = F::sub(u2, acc.x);
        const El r = F::sub(s2, acc.y);
        if (F::is_zero(p)) {
            if (F::is_zero(r)) return dbl_affine<M>(q);  // same point
            return x_inf();                              // opposite points
        }
        const El pp = M::sqr(p);
        const El ppp = M::mul(p, pp);
        const El qq = M::mul(acc.x, pp);
        X o;
        o.x = F::sub(F::sub(F::sub(M::sqr(r), ppp), qq), qq);
        o.y = F::sub(M::mul(r, F::sub(qq, o.x)), M::mul(acc.y, ppp));
        o.zz = M::mul(acc.zz, pp);
        o.zzz = M::mul(acc.zzz, ppp);
        return o;
    }
    // a + b, add-2008-s
    template <class M>
    static EC_HD X add(const X& a, const X& b) {
        if (x_is_inf(b)) return a;
        if (x_is_inf(a)) return b;
        const El u1 = M::mul(a.x, b.zz);
        const El u2 = M::mul(b.x, a.zz);
        const El s1 = M::mul(a.y, b.zzz);
        const El s2 = M::mul(b.y, a.zzz);
        const El p = F::sub(u2, u1);
        const El r = F::sub(s2, s1);
        if (F::is_zero(p)) {
            if (F::is_zero(r)) return dbl<M>(a);
            return x_inf();
        }
        const El pp = M::sqr(p);
        const El ppp = M::mul(p, pp);
        const El qq = M::mul(u1, pp);
        X o;
        o.x = F::sub(F::sub(F::sub(M::sqr(r), ppp), qq), qq);
        o.y = F::sub(M::mul(r, F::sub(qq, o.x)), M::mul(s1, ppp));
        o.zz = M::mul(M::mul(a.zz, b.zz), pp);
        o.zzz = M::mul(M::mul(a.zzz, b.zzz), ppp);
        return o;
    }
    // k * P for a small k (the bucket-chunk offsets of the window reduction), double-and-add from the top bit
    template <class M>
    static EC_HD X mul_small(const X& p, uint32_t k) {
        X acc = x_inf();
        int top = -1;
        for (int i = 0; i < 32; i++)
            if ((k >> i) & 1) top = i;
        for (int i = top; i >= 0; i--) {
            acc = dbl<M>(acc);
            if ((k >> i) & 1) acc = add<M>(acc, p);
        }
        return acc;
    }
    // affine form: x = X/ZZ = X * ZZ^2 / ZZZ^2, y = Y/ZZZ (one inversion)
    static EC_HD Affine to_affine(const X& p) {
        Affine r;
        if (x_is_inf(p)) {
            r.x = F::zero(), r.y = F::zero();
            return r;
        }
        const El a = F::inv(p.zzz);
        const El a2 = MCal::mul(a, a);
        const El zz2 = MCal::mul(p.zz, p.zz);
        r.x = MCal::mul(MCal::mul(p.x, zz2), a2);
        r.y = MCal::mul(p.y, a);
        return r;
    }
};

typedef Curve<FpBase, FpMulInline, FpMulCall> G1;
typedef Curve<Fp2Base, Fp2Mul<FpMulInline>, Fp2Mul<FpMulCall>> G2;

}  // namespace ec
