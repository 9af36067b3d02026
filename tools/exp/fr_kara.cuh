// REJECTED EXPERIMENT (round 1), kept for the record: slower than the operand-scanning multiplier on B200 (profiles/r1_exp_karatsuba.txt).
// Not part of the product build.
// BN254 Fr multiplier, Karatsuba form: 48 + 72 wide multiply-adds per Montgomery product instead of 64 + 72,
// and 48 instead of 64 for the plain 512-bit products the round sums accumulate.
//
// The integer multiplier pipe is the bound of every hot kernel (IMAD.WIDE.U32: one warp instruction per ~4.4 cycles per
// SM sub-partition) while the ALU pipe runs at ~20 %, so one level of SUBTRACTIVE Karatsuba trades
// 16 wide multiply-adds per product for ~100 carry-chain adds:
//     a = a1*B + a0, b = b1*B + b0 (B = 2^128):   a*b = z0 + (z0 + z2 + (a0 - a1)*(b1 - b0))*B + z2*B^2
// with z0 = a0*b0, z2 = a1*b1 and the middle product taken on |a0 - a1|, |b1 - b0| (sign applied afterwards).
// The Montgomery reduction then runs on the 512-bit product as eight rows m_i*q, exactly the rows of fr_mul
// (fr_device.cuh) without their a*b_i chains, so the result is the same canonical residue bit for bit.
//
// Everything here is __host__ __device__: tools/exp/kara_host_test.cu runs the same source on the CPU (portable
// fallbacks of the carry-chain primitives) against the host multiplier of fr_host.hpp.
#pragma once
#include "../../gkr-mimc_b200/csrc/fr_device.cuh"

namespace gkr {

#define GKR_HD __host__ __device__ __forceinline__

// (c0,c1) += x0*y ; (c2,c3) += x1*y ; carry-out added to top
GKR_HD void hd_chain2(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& top, uint32_t x0, uint32_t x1, uint32_t y) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(top)
        : "r"(x0), "r"(x1), "r"(y));
#else
    unsigned __int128 t = (unsigned __int128)((uint64_t)c0 | ((uint64_t)c1 << 32)) + (uint64_t)x0 * y;
    c0 = (uint32_t)t;
    c1 = (uint32_t)(t >> 32);
    t = (unsigned __int128)((uint64_t)c2 | ((uint64_t)c3 << 32)) + (uint64_t)x1 * y + (uint64_t)(t >> 64);
    c2 = (uint32_t)t;
    c3 = (uint32_t)(t >> 32);
    top += (uint32_t)(t >> 64);
#endif
}
// same, without a carry-out limb (the caller knows the columns cannot overflow)
GKR_HD void hd_chain2_nt(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t x0, uint32_t x1, uint32_t y) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\t"
        "madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
        "madc.lo.cc.u32 %2, %5, %6, %2;\n\t"
        "madc.hi.u32 %3, %5, %6, %3;"
        : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3)
        : "r"(x0), "r"(x1), "r"(y));
#else
    uint32_t top = 0;
    hd_chain2(c0, c1, c2, c3, top, x0, x1, y);
#endif
}
GKR_HD void hd_chain4(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& c4, uint32_t& c5, uint32_t& c6, uint32_t& c7, uint32_t& top,
                      uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
#ifdef __CUDA_ARCH__
    chain4(c0, c1, c2, c3, c4, c5, c6, c7, top, x0, x1, x2, x3, y);
#else
    uint32_t t1 = 0;
    hd_chain2(c0, c1, c2, c3, t1, x0, x1, y);
    // carry t1 continues into the next column
    unsigned __int128 t = (unsigned __int128)((uint64_t)c4 | ((uint64_t)c5 << 32)) + (uint64_t)x2 * y + t1;
    c4 = (uint32_t)t;
    c5 = (uint32_t)(t >> 32);
    t = (unsigned __int128)((uint64_t)c6 | ((uint64_t)c7 << 32)) + (uint64_t)x3 * y + (uint64_t)(t >> 64);
    c6 = (uint32_t)t;
    c7 = (uint32_t)(t >> 32);
    top += (uint32_t)(t >> 64);
#endif
}
// carry-in = carry of (d0 + d1), see chain4_cin
GKR_HD void hd_chain4_cin(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& c4, uint32_t& c5, uint32_t& c6, uint32_t& c7, uint32_t& top,
                          uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y, uint32_t d0, uint32_t d1) {
#ifdef __CUDA_ARCH__
    chain4_cin(c0, c1, c2, c3, c4, c5, c6, c7, top, x0, x1, x2, x3, y, d0, d1);
#else
    const uint64_t cin = ((uint64_t)d0 + d1) >> 32;
    unsigned __int128 t = (unsigned __int128)((uint64_t)c0 | ((uint64_t)c1 << 32)) + (uint64_t)x0 * y + cin;
    c0 = (uint32_t)t;
    c1 = (uint32_t)(t >> 32);
    t = (unsigned __int128)((uint64_t)c2 | ((uint64_t)c3 << 32)) + (uint64_t)x1 * y + (uint64_t)(t >> 64);
    c2 = (uint32_t)t;
    c3 = (uint32_t)(t >> 32);
    t = (unsigned __int128)((uint64_t)c4 | ((uint64_t)c5 << 32)) + (uint64_t)x2 * y + (uint64_t)(t >> 64);
    c4 = (uint32_t)t;
    c5 = (uint32_t)(t >> 32);
    t = (unsigned __int128)((uint64_t)c6 | ((uint64_t)c7 << 32)) + (uint64_t)x3 * y + (uint64_t)(t >> 64);
    c6 = (uint32_t)t;
    c7 = (uint32_t)(t >> 32);
    top += (uint32_t)(t >> 64);
#endif
}

// r[0..N) = a + b + cin (cin in {0,1}); returns the carry out (0/1).  One carry chain inside ONE asm statement per length.
GKR_HD uint32_t hd_add4(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
    uint32_t c;
    asm("add.cc.u32 %0, %5, %9;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.cc.u32 %3, %8, %12;\n\t"
        "addc.u32 %4, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
    return c;
#else
    uint64_t c = 0;
    for (int i = 0; i < 4; i++) {
        c += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
#endif
}
// r = a - b over 4 limbs; returns the borrow as a mask (0xffffffff when a < b)
GKR_HD uint32_t hd_sub4(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
    uint32_t m;
    asm("sub.cc.u32 %0, %5, %9;\n\t"
        "subc.cc.u32 %1, %6, %10;\n\t"
        "subc.cc.u32 %2, %7, %11;\n\t"
        "subc.cc.u32 %3, %8, %12;\n\t"
        "subc.u32 %4, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(m)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
    return m;
#else
    int64_t c = 0;
    for (int i = 0; i < 4; i++) {
        c += (int64_t)a[i] - (int64_t)b[i];
        r[i] = (uint32_t)c;
        c >>= 32;  // arithmetic shift: 0 or -1
    }
    return (uint32_t)c;
#endif
}
// x = (x ^ m) - m over 4 limbs (m = 0: unchanged; m = all ones: two's complement negation)
GKR_HD void hd_cneg4(uint32_t* x, uint32_t m) {
#ifdef __CUDA_ARCH__
    asm("sub.cc.u32 %0, %0, %4;\n\t"
        "subc.cc.u32 %1, %1, %4;\n\t"
        "subc.cc.u32 %2, %2, %4;\n\t"
        "subc.u32 %3, %3, %4;"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3])
        : "r"(m));
    // note: the xor is applied by the caller BEFORE this call (keeps the asm to one carry chain)
#else
    int64_t c = 0;
    for (int i = 0; i < 4; i++) {
        c += (int64_t)x[i] - (int64_t)m;
        x[i] = (uint32_t)c;
        c >>= 32;
    }
#endif
}

// z[0..8) = x[0..4) * y[0..4): 16 wide multiply-adds in the two-accumulator parity scheme of fr_mul (P: 64-bit columns at even
// limb positions, Q: at odd positions), eight independent 2-product chains.
GKR_HD void hd_mul4(const uint32_t* x, const uint32_t* y, uint32_t* z) {
    uint32_t P[9], Q[9];
#pragma unroll
    for (int i = 0; i < 9; i++) P[i] = 0, Q[i] = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t* S = (i & 1) ? Q : P;
        uint32_t* T = (i & 1) ? P : Q;
        if (i < 3) {
            hd_chain2(S[i], S[i + 1], S[i + 2], S[i + 3], S[i + 4], x[0], x[2], y[i]);
            hd_chain2(T[i + 1], T[i + 2], T[i + 3], T[i + 4], T[i + 5], x[1], x[3], y[i]);
        } else {
            // last row: each accumulator stays below the full product < 2^256, nothing can leave limb 7
            hd_chain2(S[i], S[i + 1], S[i + 2], S[i + 3], S[i + 4], x[0], x[2], y[i]);
            hd_chain2_nt(T[i + 1], T[i + 2], T[i + 3], T[i + 4], x[1], x[3], y[i]);
        }
    }
    // z = P + Q (< 2^256: no carry out of limb 7)
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(z[0]), "=r"(z[1]), "=r"(z[2]), "=r"(z[3]), "=r"(z[4]), "=r"(z[5]), "=r"(z[6]), "=r"(z[7])
        : "r"(P[0]), "r"(P[1]), "r"(P[2]), "r"(P[3]), "r"(P[4]), "r"(P[5]), "r"(P[6]), "r"(P[7]),
          "r"(Q[0]), "r"(Q[1]), "r"(Q[2]), "r"(Q[3]), "r"(Q[4]), "r"(Q[5]), "r"(Q[6]), "r"(Q[7]));
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)P[i] + Q[i];
        z[i] = (uint32_t)c;
        c >>= 32;
    }
#endif
}

// t[0..16) = a * b (plain 512-bit product of two 256-bit values), 48 wide multiply-adds
GKR_HD void hd_mul_wide_k(const Fr& a, const Fr& b, uint32_t* t) {
    uint32_t z0[8], z2[8], zm[8], da[4], db[4];
    hd_mul4(a.v, b.v, z0);
    hd_mul4(a.v + 4, b.v + 4, z2);
    // da = |a0 - a1|, db = |b1 - b0|, s = sign of (a0 - a1)*(b1 - b0) as a mask
    const uint32_t sa = hd_sub4(da, a.v, a.v + 4);
    const uint32_t sb = hd_sub4(db, b.v + 4, b.v);
#pragma unroll
    for (int i = 0; i < 4; i++) da[i] ^= sa, db[i] ^= sb;
    hd_cneg4(da, sa);
    hd_cneg4(db, sb);
    hd_mul4(da, db, zm);
    const uint32_t s = sa ^ sb;
    // mid = z0 + z2 + (s ? -zm : zm), 9 limbs, always >= 0 (it equals a0*b1 + a1*b0)
    uint32_t mid[9];
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(mid[0]), "=r"(mid[1]), "=r"(mid[2]), "=r"(mid[3]), "=r"(mid[4]), "=r"(mid[5]), "=r"(mid[6]), "=r"(mid[7]), "=r"(mid[8])
        : "r"(z0[0]), "r"(z0[1]), "r"(z0[2]), "r"(z0[3]), "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]),
          "r"(z2[0]), "r"(z2[1]), "r"(z2[2]), "r"(z2[3]), "r"(z2[4]), "r"(z2[5]), "r"(z2[6]), "r"(z2[7]));
#pragma unroll
    for (int i = 0; i < 8; i++) zm[i] ^= s;
    // + (zm ^ s) + (s & 1) with the sign extension s in limb 8: two's complement addition, wrap-around carry dropped
    uint32_t scratch;
    asm("add.cc.u32 %9, %10, %10;\n\t"   // sets the carry flag to (s != 0): s + s overflows exactly when s = 0xffffffff
        "addc.cc.u32 %0, %0, %11;\n\t"
        "addc.cc.u32 %1, %1, %12;\n\t"
        "addc.cc.u32 %2, %2, %13;\n\t"
        "addc.cc.u32 %3, %3, %14;\n\t"
        "addc.cc.u32 %4, %4, %15;\n\t"
        "addc.cc.u32 %5, %5, %16;\n\t"
        "addc.cc.u32 %6, %6, %17;\n\t"
        "addc.cc.u32 %7, %7, %18;\n\t"
        "addc.u32 %8, %8, %10;"
        : "+r"(mid[0]), "+r"(mid[1]), "+r"(mid[2]), "+r"(mid[3]), "+r"(mid[4]), "+r"(mid[5]), "+r"(mid[6]), "+r"(mid[7]), "+r"(mid[8]), "=r"(scratch)
        : "r"(s), "r"(zm[0]), "r"(zm[1]), "r"(zm[2]), "r"(zm[3]), "r"(zm[4]), "r"(zm[5]), "r"(zm[6]), "r"(zm[7]));
    // t = z0 + mid * 2^128 + z2 * 2^256
    t[0] = z0[0], t[1] = z0[1], t[2] = z0[2], t[3] = z0[3];
    uint32_t c1;
    asm("add.cc.u32 %0, %5, %9;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.cc.u32 %3, %8, %12;\n\t"
        "addc.u32 %4, 0, 0;"
        : "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(c1)
        : "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]), "r"(mid[0]), "r"(mid[1]), "r"(mid[2]), "r"(mid[3]));
    asm("add.cc.u32 %8, %9, 0xffffffff;\n\t"   // carry flag = c1
        "addc.cc.u32 %0, %10, %18;\n\t"
        "addc.cc.u32 %1, %11, %19;\n\t"
        "addc.cc.u32 %2, %12, %20;\n\t"
        "addc.cc.u32 %3, %13, %21;\n\t"
        "addc.cc.u32 %4, %14, %22;\n\t"
        "addc.cc.u32 %5, %15, 0;\n\t"
        "addc.cc.u32 %6, %16, 0;\n\t"
        "addc.u32 %7, %17, 0;"
        : "=r"(t[8]), "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15]), "=r"(scratch)
        : "r"(c1), "r"(z2[0]), "r"(z2[1]), "r"(z2[2]), "r"(z2[3]), "r"(z2[4]), "r"(z2[5]), "r"(z2[6]), "r"(z2[7]), "r"(mid[4]), "r"(mid[5]), "r"(mid[6]),
          "r"(mid[7]), "r"(mid[8]));
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)z0[i] + z2[i];
        mid[i] = (uint32_t)c;
        c >>= 32;
    }
    mid[8] = (uint32_t)c;
    c = s & 1;
    for (int i = 0; i < 9; i++) {
        c += (uint64_t)mid[i] + (i < 8 ? (zm[i] ^ s) : s);
        mid[i] = (uint32_t)c;
        c >>= 32;
    }
    for (int i = 0; i < 4; i++) t[i] = z0[i];
    c = 0;
    for (int i = 4; i < 16; i++) {
        c += (uint64_t)(i < 8 ? z0[i] : z2[i - 8]) + (i < 13 ? mid[i - 4] : 0);
        t[i] = (uint32_t)c;
        c >>= 32;
    }
#endif
}

// Montgomery reduction of a 512-bit value t < q * 2^256: t * 2^-256 mod q, canonical.  The eight rows of fr_mul without
// their a*b_i chains (72 wide multiply-adds).
GKR_HD Fr hd_redc(const uint32_t* t) {
    // The rows run on t mod 2^256 only: the limbs a chain uses as its carry-out (`top`) then hold nothing but earlier carries
    // and cannot wrap (with the high half of t in place a limb equal to 0xffffffff would).  t >> 256 is added at the end.
    uint32_t P[18], Qd[18];
#pragma unroll
    for (int i = 0; i < 18; i++) P[i] = i < 8 ? t[i] : 0, Qd[i] = 0;
    const uint32_t q[8] = {FR_Q0, FR_Q1, FR_Q2, FR_Q3, FR_Q4, FR_Q5, FR_Q6, FR_Q7};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t* S = (i & 1) ? Qd : P;
        uint32_t* T = (i & 1) ? P : Qd;
        const uint32_t m = (S[i] + T[i]) * FR_QINV32;
        hd_chain4(S[i], S[i + 1], S[i + 2], S[i + 3], S[i + 4], S[i + 5], S[i + 6], S[i + 7], S[i + 8], q[0], q[2], q[4], q[6], m);
        hd_chain4_cin(T[i + 1], T[i + 2], T[i + 3], T[i + 4], T[i + 5], T[i + 6], T[i + 7], T[i + 8], T[i + 9], q[1], q[3], q[5], q[7], m, S[i], T[i]);
    }
    Fr r;
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(P[8]), "r"(P[9]), "r"(P[10]), "r"(P[11]), "r"(P[12]), "r"(P[13]), "r"(P[14]), "r"(P[15]),
          "r"(Qd[8]), "r"(Qd[9]), "r"(Qd[10]), "r"(Qd[11]), "r"(Qd[12]), "r"(Qd[13]), "r"(Qd[14]), "r"(Qd[15]));
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7])
        : "r"(t[8]), "r"(t[9]), "r"(t[10]), "r"(t[11]), "r"(t[12]), "r"(t[13]), "r"(t[14]), "r"(t[15]));
    return fr_reduce_once(r);  // (t + M*q) / 2^256 < 2q
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)P[8 + i] + Qd[8 + i] + t[8 + i];
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    // canonicalise (< 2q)
    const uint32_t qq[8] = {FR_Q0, FR_Q1, FR_Q2, FR_Q3, FR_Q4, FR_Q5, FR_Q6, FR_Q7};
    Fr s;
    int64_t bw = 0;
    for (int i = 0; i < 8; i++) {
        bw += (int64_t)r.v[i] - (int64_t)qq[i];
        s.v[i] = (uint32_t)bw;
        bw >>= 32;
    }
    return bw ? r : s;
#endif
}

// Montgomery product, Karatsuba form: same result as fr_mul, 120 wide multiply-adds instead of 136
GKR_HD Fr hd_mul_k(const Fr& a, const Fr& b) {
    uint32_t t[16];
    hd_mul_wide_k(a, b, t);
    return hd_redc(t);
}

}  // namespace gkr
