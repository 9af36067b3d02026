// BN254 Fr on the device: Montgomery form, 8 x 32-bit limbs in registers, sm_100a.
//
// Replaces gnark-crypto's fr.Element (reference go.mod:7, package ecc/bn254/fr; amd64 asm in the
// un-vendored module) for the GPU kernels.  Memory format is unchanged from Go: 4 x u64 little-endian
// limbs, value*2^256 mod q, canonical (< q); a 32-byte aligned element is moved with ONE 256-bit
// LDG/STG (sm_100 `ld.global.v4.u64`), so a warp touches 1 KiB contiguous.
//
// Multiplication: operand-scanning Montgomery product on IMAD.WIDE.U32 with carry chaining
// (mad.lo.cc/madc.hi.cc pairs fuse into IMAD.WIDE.U32(.X)).  64-bit partial products land in two
// accumulators split by the PARITY OF THEIR ABSOLUTE LIMB POSITION, so every product is added with a
// single 64-bit-aligned IMAD.WIDE and no limb shuffling; a row's dead low limb is folded into the next
// chain as its carry-in.  136 wide multiplies per product (64 a*b + 64 m*q + 8 m), see DESIGN.md.
#pragma once
#include <cstdint>

namespace gkr {

struct __align__(32) FrRaw {  // memory image == Go fr.Element
    uint64_t l[4];
};

struct Fr {
    uint32_t v[8];
};

// q and friends as 32-bit limbs (re-derived in tests/test_oracle.py)
#define FR_Q0 0xf0000001u
#define FR_Q1 0x43e1f593u
#define FR_Q2 0x79b97091u
#define FR_Q3 0x2833e848u
#define FR_Q4 0x8181585du
#define FR_Q5 0xb85045b6u
#define FR_Q6 0xe131a029u
#define FR_Q7 0x30644e72u
#define FR_QINV32 0xefffffffu  // -q^{-1} mod 2^32

__device__ __forceinline__ Fr fr_zero() {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
__device__ __forceinline__ Fr fr_one() {  // 2^256 mod q
    Fr r;
    r.v[0] = 0x4ffffffbu; r.v[1] = 0xac96341cu; r.v[2] = 0x9f60cd29u; r.v[3] = 0x36fc7695u;
    r.v[4] = 0x7879462eu; r.v[5] = 0x666ea36fu; r.v[6] = 0x9a07df2fu; r.v[7] = 0x0e0a77c1u;
    return r;
}

__device__ __forceinline__ Fr fr_unpack(const FrRaw& m) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        r.v[2 * i] = (uint32_t)m.l[i];
        r.v[2 * i + 1] = (uint32_t)(m.l[i] >> 32);
    }
    return r;
}
__device__ __forceinline__ FrRaw fr_pack(const Fr& a) {
    FrRaw m;
#pragma unroll
    for (int i = 0; i < 4; i++) m.l[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
    return m;
}
// 256-bit global accesses (LDG.E.256 / STG.E.256 on sm_100)
__device__ __forceinline__ Fr fr_load(const FrRaw* p) {
    FrRaw m;
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(m.l[0]), "=l"(m.l[1]), "=l"(m.l[2]), "=l"(m.l[3]) : "l"(p));
    return fr_unpack(m);
}
// read-only / streaming variant: data read once, do not keep in L1
__device__ __forceinline__ Fr fr_load_stream(const FrRaw* p) {
    FrRaw m;
    asm volatile("ld.global.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(m.l[0]), "=l"(m.l[1]), "=l"(m.l[2]), "=l"(m.l[3]) : "l"(p));
    return fr_unpack(m);
}
__device__ __forceinline__ void fr_store(FrRaw* p, const Fr& a) {
    FrRaw m = fr_pack(a);
    asm volatile("st.global.v4.u64 [%4], {%0,%1,%2,%3};" ::"l"(m.l[0]), "l"(m.l[1]), "l"(m.l[2]), "l"(m.l[3]), "l"(p) : "memory");
}

// r = a - q, returns borrow (1 when a < q)
__device__ __forceinline__ uint32_t fr_sub_q(Fr& r, const Fr& a) {
    uint32_t borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]), "=r"(borrow)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "n"(FR_Q0), "n"(FR_Q1), "n"(FR_Q2), "n"(FR_Q3), "n"(FR_Q4), "n"(FR_Q5), "n"(FR_Q6), "n"(FR_Q7));
    return borrow;  // 0xffffffff when a < q, else 0
}
// canonicalise a value known to be < 2q
__device__ __forceinline__ Fr fr_reduce_once(const Fr& a) {
    Fr s;
    uint32_t borrow = fr_sub_q(s, a);
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = borrow ? a.v[i] : s.v[i];
    return r;
}

// fr.Element.Add
__device__ __forceinline__ Fr fr_add(const Fr& a, const Fr& b) {
    Fr t;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(t.v[0]), "=r"(t.v[1]), "=r"(t.v[2]), "=r"(t.v[3]), "=r"(t.v[4]), "=r"(t.v[5]), "=r"(t.v[6]), "=r"(t.v[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    return fr_reduce_once(t);  // a+b < 2q < 2^255: no carry out of limb 7
}

// fr.Element.Sub
__device__ __forceinline__ Fr fr_sub(const Fr& a, const Fr& b) {
    Fr t;
    uint32_t borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(t.v[0]), "=r"(t.v[1]), "=r"(t.v[2]), "=r"(t.v[3]), "=r"(t.v[4]), "=r"(t.v[5]), "=r"(t.v[6]), "=r"(t.v[7]), "=r"(borrow)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    // add back q masked by the borrow
    Fr r;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(t.v[0]), "r"(t.v[1]), "r"(t.v[2]), "r"(t.v[3]), "r"(t.v[4]), "r"(t.v[5]), "r"(t.v[6]), "r"(t.v[7]),
          "r"(borrow & FR_Q0), "r"(borrow & FR_Q1), "r"(borrow & FR_Q2), "r"(borrow & FR_Q3), "r"(borrow & FR_Q4), "r"(borrow & FR_Q5),
          "r"(borrow & FR_Q6), "r"(borrow & FR_Q7));
    return r;
}
__device__ __forceinline__ Fr fr_dbl(const Fr& a) { return fr_add(a, a); }
__device__ __forceinline__ Fr fr_neg(const Fr& a) { return fr_sub(fr_zero(), a); }
__device__ __forceinline__ bool fr_is_zero(const Fr& a) {
    return (a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7]) == 0;
}

// ---------------------------------------------------------------------------------------------
// Carry-chained multiply-accumulate of four 64-bit columns:
//   (c0,c1) += x0*y ; (c2,c3) += x1*y ; (c4,c5) += x2*y ; (c6,c7) += x3*y   (+ carry-in), carry-out added to `top`.
// Each mad.lo.cc/madc.hi.cc pair is one IMAD.WIDE.U32(.X).  The whole chain lives in ONE asm
// statement so the condition-code register never crosses a statement boundary.
// ---------------------------------------------------------------------------------------------
#define GKR_CHAIN_BODY                      \
    "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"   \
    "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"  \
    "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"  \
    "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"  \
    "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"  \
    "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"  \
    "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"  \
    "addc.u32 %8, %8, 0;"

// no carry-in
__device__ __forceinline__ void chain4(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& c4, uint32_t& c5, uint32_t& c6,
                                       uint32_t& c7, uint32_t& top, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t" GKR_CHAIN_BODY
        : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(c6), "+r"(c7), "+r"(top)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
}
// carry-in = carry of (d0 + d1): the dead low limb of the running sum (d0 + d1 == 0 mod 2^32)
__device__ __forceinline__ void chain4_cin(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& c4, uint32_t& c5, uint32_t& c6,
                                           uint32_t& c7, uint32_t& top, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y,
                                           uint32_t d0, uint32_t d1) {
    uint32_t dead;
    asm("add.cc.u32 %9, %15, %16;\n\t"
        "madc.lo.cc.u32 %0, %10, %14, %0;\n\t"
        "madc.hi.cc.u32 %1, %10, %14, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %14, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %14, %6;\n\t"
        "madc.hi.cc.u32 %7, %13, %14, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(c6), "+r"(c7), "+r"(top), "=r"(dead)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y), "r"(d0), "r"(d1));
    (void)dead;
}

// m = x * (-q^-1) mod 2^32.  -q^-1 = 0xefffffff = -(2^28 + 1), so m = -(x + (x << 28)): a shift-add and a negation on the
// ALU pipe instead of one more IMAD on the pipe the wide multiply-adds are bound by (8 per product).
#ifndef GKR_M_BY_SHIFT
#define GKR_M_BY_SHIFT 1
#endif
__device__ __forceinline__ uint32_t fr_mont_m(uint32_t x) {
#if GKR_M_BY_SHIFT
    return 0u - (x + (x << 28));
#else
    return x * FR_QINV32;
#endif
}

// Montgomery product a*b*2^-256 mod q, inputs canonical (< q), output canonical.
//
// P[p] is the accumulator limb at ABSOLUTE position p (weight 2^(32p)) of the accumulator whose
// 64-bit columns start at even positions; Qd[p] the same for odd-aligned columns.  Row i adds
// a*b_i*2^(32i) and m_i*q*2^(32i), which zeroes limb i of P+Qd (mod 2^32); its carry rides into the next chain.
__device__ __forceinline__ Fr fr_mul_school(const Fr& a, const Fr& b) {
    uint32_t P[18], Qd[18];
#pragma unroll
    for (int i = 0; i < 18; i++) P[i] = 0, Qd[i] = 0;
    const uint32_t q[8] = {FR_Q0, FR_Q1, FR_Q2, FR_Q3, FR_Q4, FR_Q5, FR_Q6, FR_Q7};

#pragma unroll
    for (int i = 0; i < 8; i++) {
        // S = accumulator whose columns start at position i (same parity as i); T = the other one
        uint32_t* S = (i & 1) ? Qd : P;
        uint32_t* T = (i & 1) ? P : Qd;
        const uint32_t bi = b.v[i];
        // columns (i,i+1)..(i+6,i+7) += a_even * b_i
        chain4(S[i], S[i + 1], S[i + 2], S[i + 3], S[i + 4], S[i + 5], S[i + 6], S[i + 7], S[i + 8], a.v[0], a.v[2], a.v[4], a.v[6], bi);
        // m makes limb i of (S + T) vanish
        const uint32_t m = fr_mont_m(S[i] + T[i]);
        chain4(S[i], S[i + 1], S[i + 2], S[i + 3], S[i + 4], S[i + 5], S[i + 6], S[i + 7], S[i + 8], q[0], q[2], q[4], q[6], m);
        // columns (i+1,i+2)..(i+7,i+8) += a_odd * b_i, carry-in = carry of the dead limb i
        chain4_cin(T[i + 1], T[i + 2], T[i + 3], T[i + 4], T[i + 5], T[i + 6], T[i + 7], T[i + 8], T[i + 9], a.v[1], a.v[3], a.v[5], a.v[7],
                   bi, S[i], T[i]);
        chain4(T[i + 1], T[i + 2], T[i + 3], T[i + 4], T[i + 5], T[i + 6], T[i + 7], T[i + 8], T[i + 9], q[1], q[3], q[5], q[7], m);
    }
    // result = limbs 8..15 of P + Qd  (< 2q)
    Fr t;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(t.v[0]), "=r"(t.v[1]), "=r"(t.v[2]), "=r"(t.v[3]), "=r"(t.v[4]), "=r"(t.v[5]), "=r"(t.v[6]), "=r"(t.v[7])
        : "r"(P[8]), "r"(P[9]), "r"(P[10]), "r"(P[11]), "r"(P[12]), "r"(P[13]), "r"(P[14]), "r"(P[15]),
          "r"(Qd[8]), "r"(Qd[9]), "r"(Qd[10]), "r"(Qd[11]), "r"(Qd[12]), "r"(Qd[13]), "r"(Qd[14]), "r"(Qd[15]));
    return fr_reduce_once(t);
}
__device__ __forceinline__ Fr fr_mul(const Fr& a, const Fr& b) { return fr_mul_school(a, b); }

// Out-of-line copy of the multiplier (register ABI, no stack: 16 words in, 8 out) for the small, latency-bound kernels:
// one ~3 KB body instead of 20-50 inlined copies keeps their cold instruction fetch short (profiles/: the fully inlined
// round kernels of round 1 had a 40 us floor for one-wave launches).  The big-round kernels inline the multiplier
// (GKR_INLINE_BIG, kernels.cuh): the call ABI costs ~375 IMAD.MOV per pair on the very pipe the multiplier is bound by.
__device__ __noinline__ Fr fr_mulc(const Fr a, const Fr b) { return fr_mul(a, b); }
__device__ __forceinline__ Fr fr_sqrc(const Fr& a) { return fr_mulc(a, a); }

// a[0..8) += b[0..8) (+ cin); returns the carry out (0 or 1).  The whole carry chain lives in ONE asm statement and the
// carry crosses statements in a register, never in the condition code.
__device__ __forceinline__ uint32_t add8_carry(uint32_t* a, const uint32_t* b) {
    uint32_t cout;
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, %12;\n\t"
        "addc.cc.u32 %4, %4, %13;\n\t"
        "addc.cc.u32 %5, %5, %14;\n\t"
        "addc.cc.u32 %6, %6, %15;\n\t"
        "addc.cc.u32 %7, %7, %16;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "=r"(cout)
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return cout;
}
__device__ __forceinline__ uint32_t add8_carry_in(uint32_t* a, const uint32_t* b, uint32_t cin) {
    uint32_t cout, dead;
    asm("add.cc.u32 %9, %18, 0xffffffff;\n\t"  // sets the carry iff cin == 1
        "addc.cc.u32 %0, %0, %10;\n\t"
        "addc.cc.u32 %1, %1, %11;\n\t"
        "addc.cc.u32 %2, %2, %12;\n\t"
        "addc.cc.u32 %3, %3, %13;\n\t"
        "addc.cc.u32 %4, %4, %14;\n\t"
        "addc.cc.u32 %5, %5, %15;\n\t"
        "addc.cc.u32 %6, %6, %16;\n\t"
        "addc.cc.u32 %7, %7, %17;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "=r"(cout), "=r"(dead)
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(cin));
    (void)dead;
    return cout;
}

// Plain 512-bit product a*b (no Montgomery reduction) into w[0..16): 64 wide multiply-adds.
__device__ __forceinline__ void fr_mul_wide(uint32_t (&w)[16], const Fr& a, const Fr& b) {
    uint32_t P[18], Qd[18];
#pragma unroll
    for (int i = 0; i < 18; i++) P[i] = 0, Qd[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t* S = (i & 1) ? Qd : P;
        uint32_t* T = (i & 1) ? P : Qd;
        const uint32_t bi = b.v[i];
        chain4(S[i], S[i + 1], S[i + 2], S[i + 3], S[i + 4], S[i + 5], S[i + 6], S[i + 7], S[i + 8], a.v[0], a.v[2], a.v[4], a.v[6], bi);
        chain4(T[i + 1], T[i + 2], T[i + 3], T[i + 4], T[i + 5], T[i + 6], T[i + 7], T[i + 8], T[i + 9], a.v[1], a.v[3], a.v[5], a.v[7], bi);
    }
    // product = P + Qd (< 2^512: limb 16 of the sum is zero)
    const uint32_t c = add8_carry(P, Qd);
    (void)add8_carry_in(P + 8, Qd + 8, c);
#pragma unroll
    for (int l = 0; l < 16; l++) w[l] = P[l];
}

// acc (17 x 32-bit limbs, one every `stride` words) += a*b as a PLAIN 512-bit product: no Montgomery reduction.
// Half the multiplier work of fr_mul (64 of its 136 wide multiply-adds); whoever consumes the sum reduces it once
// (the host, for the round sums: REDC(sum of products) == sum of Montgomery products, exactly).  544 bits hold
// 2^36 products of values < q.
__device__ __forceinline__ void fr_mul_acc_wide_inl(uint32_t* acc, int stride, const Fr& a, const Fr& b) {
    uint32_t p[16];
    fr_mul_wide(p, a, b);
    uint32_t w[17];
#pragma unroll
    for (int l = 0; l < 17; l++) w[l] = acc[l * stride];
    uint32_t c = add8_carry(w, p);
    c = add8_carry_in(w + 8, p + 8, c);
    w[16] += c;
#pragma unroll
    for (int l = 0; l < 17; l++) acc[l * stride] = w[l];
}
__device__ __noinline__ void fr_mul_acc_wide(uint32_t* acc, int stride, const Fr a, const Fr b) { fr_mul_acc_wide_inl(acc, stride, a, b); }

// Montgomery reduction of a 512-bit value t < q * 2^256: t * 2^-256 mod q, canonical.  The eight rows of fr_mul without
// their a*b_i chains (72 wide multiply-adds).  The rows run on t mod 2^256 only: the limbs a chain uses as its carry-out
// (`top`) then hold nothing but earlier carries and cannot wrap; t >> 256 is added at the end.
__device__ __forceinline__ Fr fr_redc_wide(const uint32_t (&t)[16]) {
    uint32_t P[18], Qd[18];
#pragma unroll
    for (int i = 0; i < 18; i++) P[i] = i < 8 ? t[i] : 0, Qd[i] = 0;
    const uint32_t q[8] = {FR_Q0, FR_Q1, FR_Q2, FR_Q3, FR_Q4, FR_Q5, FR_Q6, FR_Q7};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t* S = (i & 1) ? Qd : P;
        uint32_t* T = (i & 1) ? P : Qd;
        const uint32_t m = fr_mont_m(S[i] + T[i]);
        chain4(S[i], S[i + 1], S[i + 2], S[i + 3], S[i + 4], S[i + 5], S[i + 6], S[i + 7], S[i + 8], q[0], q[2], q[4], q[6], m);
        chain4_cin(T[i + 1], T[i + 2], T[i + 3], T[i + 4], T[i + 5], T[i + 6], T[i + 7], T[i + 8], T[i + 9], q[1], q[3], q[5], q[7], m, S[i], T[i]);
    }
    uint32_t r[8], hi[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = P[8 + i], hi[i] = t[8 + i];
    (void)add8_carry(r, Qd + 8);
    (void)add8_carry(r, hi);  // (t + M*q) / 2^256 < 2q < 2^255: no carry out
    Fr v;
#pragma unroll
    for (int i = 0; i < 8; i++) v.v[i] = r[i];
    return fr_reduce_once(v);
}

// ---------------------------------------------------------------------------------------------
// Dedicated squaring: 36 wide multiply-adds for the 512-bit square instead of 64 (28 off-diagonal products a_i*a_j, i < j,
// taken once and doubled with 16 funnel shifts on the ALU pipe, + 8 diagonal squares), then the 72 of the Montgomery
// reduction: 108 instead of 136 per fr.Element.Square.  Same canonical result as fr_mul(a, a) (tests compare both).
// Chains of 1-3 columns for the ragged rows; as in fr_mul_school the limb a chain carries out into (`top`) holds nothing
// but earlier carries when the chain runs (rows in ascending i), so it cannot wrap.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void chain1(uint32_t& c0, uint32_t& c1, uint32_t& top, uint32_t x0, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(c0), "+r"(c1), "+r"(top)
        : "r"(x0), "r"(y));
}
__device__ __forceinline__ void chain2(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& top, uint32_t x0, uint32_t x1, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(top)
        : "r"(x0), "r"(x1), "r"(y));
}
__device__ __forceinline__ void chain3(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& c4, uint32_t& c5, uint32_t& top, uint32_t x0,
                                       uint32_t x1, uint32_t x2, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
        "addc.u32 %6, %6, 0;"
        : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(top)
        : "r"(x0), "r"(x1), "r"(x2), "r"(y));
}
// w[0..16) = a^2 as a plain 512-bit integer
__device__ __forceinline__ void fr_sqr_wide(uint32_t (&w)[16], const Fr& x) {
    const uint32_t* a = x.v;
    uint32_t P[16], Qd[16];  // off-diagonal sum, columns at even / odd positions
#pragma unroll
    for (int i = 0; i < 16; i++) P[i] = 0, Qd[i] = 0;
    // row i: a_i * a_j for j > i; j - i even -> position i + j even -> P, odd -> Qd
    chain3(P[2], P[3], P[4], P[5], P[6], P[7], P[8], a[2], a[4], a[6], a[0]);
    chain4(Qd[1], Qd[2], Qd[3], Qd[4], Qd[5], Qd[6], Qd[7], Qd[8], Qd[9], a[1], a[3], a[5], a[7], a[0]);
    chain3(P[4], P[5], P[6], P[7], P[8], P[9], P[10], a[3], a[5], a[7], a[1]);
    chain3(Qd[3], Qd[4], Qd[5], Qd[6], Qd[7], Qd[8], Qd[9], a[2], a[4], a[6], a[1]);
    chain2(P[6], P[7], P[8], P[9], P[10], a[4], a[6], a[2]);
    chain3(Qd[5], Qd[6], Qd[7], Qd[8], Qd[9], Qd[10], Qd[11], a[3], a[5], a[7], a[2]);
    chain2(P[8], P[9], P[10], P[11], P[12], a[5], a[7], a[3]);
    chain2(Qd[7], Qd[8], Qd[9], Qd[10], Qd[11], a[4], a[6], a[3]);
    chain1(P[10], P[11], P[12], a[6], a[4]);
    chain2(Qd[9], Qd[10], Qd[11], Qd[12], Qd[13], a[5], a[7], a[4]);
    chain1(P[12], P[13], P[14], a[7], a[5]);
    chain1(Qd[11], Qd[12], Qd[13], a[6], a[5]);
    chain1(Qd[13], Qd[14], Qd[15], a[7], a[6]);
    // S = P + Qd (< 2^481), doubled
    const uint32_t c = add8_carry(P, Qd);
    (void)add8_carry_in(P + 8, Qd + 8, c);
    w[0] = P[0] << 1;
#pragma unroll
    for (int l = 1; l < 16; l++) w[l] = __funnelshift_l(P[l - 1], P[l], 1);
    // + the diagonal a_i^2 at position 2i: one 16-limb carry chain (the total is a^2 < 2^512: no carry out)
    asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\t"
        "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
        "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
        "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
        "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
        "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
        "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
        "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
        "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
        "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
        "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
        "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
        "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
        "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
        "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
        "madc.hi.u32 %15, %23, %23, %15;"
        : "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]), "+r"(w[8]), "+r"(w[9]), "+r"(w[10]),
          "+r"(w[11]), "+r"(w[12]), "+r"(w[13]), "+r"(w[14]), "+r"(w[15])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
}
// fr.Element.Square
#ifndef GKR_FAST_SQR
#define GKR_FAST_SQR 1
#endif
__device__ __forceinline__ Fr fr_sqr(const Fr& a) {
#if GKR_FAST_SQR
    uint32_t w[16];
    fr_sqr_wide(w, a);
    return fr_redc_wide(w);
#else
    return fr_mul(a, a);
#endif
}

// ---------------------------------------------------------------------------------------------
// Product by a per-launch constant (the fold challenge r: every fold of a sumcheck round is r * (top - bottom)).
// With K_i = r * 2^(32i+64) * 2^-256 mod q (eight canonical residues the host derives from r once per launch),
//     sum_i d_i * K_i  ==  (r * d * 2^-256) * 2^64   (mod q),   d = sum_i d_i 2^(32i),
// a 290-bit number built from 64 wide multiply-adds whose rows all start at limb 0; two Montgomery rows remove the factor
// 2^64: 80 wide multiply-adds instead of 136, same canonical result as fr_mul(r, d).  K lives in the kernel's parameter
// space, so its limbs reach the multiplier as constant-bank operands and cost no registers.
// ---------------------------------------------------------------------------------------------
struct FrConstMul {
    uint32_t k[8][8];
};
__device__ __forceinline__ Fr fr_mul_const(const FrConstMul& K, const Fr& d) {
    uint32_t P[11], Qd[10];
#pragma unroll
    for (int i = 0; i < 11; i++) P[i] = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) Qd[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        chain4(P[0], P[1], P[2], P[3], P[4], P[5], P[6], P[7], P[8], K.k[i][0], K.k[i][2], K.k[i][4], K.k[i][6], d.v[i]);
        chain4(Qd[1], Qd[2], Qd[3], Qd[4], Qd[5], Qd[6], Qd[7], Qd[8], Qd[9], K.k[i][1], K.k[i][3], K.k[i][5], K.k[i][7], d.v[i]);
    }
    // the sum is < 2^35 * q: P[8] and Qd[9] hold a few carries, the two reduction rows below add < 2^30 to limb 9
    uint32_t m = fr_mont_m(P[0] + Qd[0]);
    chain4(P[0], P[1], P[2], P[3], P[4], P[5], P[6], P[7], P[8], FR_Q0, FR_Q2, FR_Q4, FR_Q6, m);
    chain4_cin(Qd[1], Qd[2], Qd[3], Qd[4], Qd[5], Qd[6], Qd[7], Qd[8], Qd[9], FR_Q1, FR_Q3, FR_Q5, FR_Q7, m, P[0], Qd[0]);
    m = fr_mont_m(Qd[1] + P[1]);
    chain4(Qd[1], Qd[2], Qd[3], Qd[4], Qd[5], Qd[6], Qd[7], Qd[8], Qd[9], FR_Q0, FR_Q2, FR_Q4, FR_Q6, m);
    chain4_cin(P[2], P[3], P[4], P[5], P[6], P[7], P[8], P[9], P[10], FR_Q1, FR_Q3, FR_Q5, FR_Q7, m, Qd[1], P[1]);
    (void)add8_carry(P + 2, Qd + 2);  // limbs 2..9 of P + Qd: the result, < 2q < 2^255
    Fr v;
#pragma unroll
    for (int i = 0; i < 8; i++) v.v[i] = P[2 + i];
    return fr_reduce_once(v);
}

// x^7 = ((x^2 * x)^2) * x  -- same chain as hash/poseidon.go:129-135 and circuit/gates/cipher.go:37-40
__device__ __forceinline__ Fr fr_pow7(const Fr& x) {
    Fr t = fr_sqr(x);
    t = fr_mul(t, x);
    t = fr_sqr(t);
    return fr_mul(t, x);
}

}  // namespace gkr
