// 254-bit Montgomery fields for the elliptic-curve path (BN254 base field Fp for G1 coordinates, scalar field Fr for the
// Montgomery -> regular conversion of the scalars): 8 x 32-bit limbs, generic over the modulus.
//
// Replaces gnark-crypto's ecc/bn254/fp.Element (reference go.mod:7, un-vendored amd64 assembly) for the GPU kernels; memory
// format unchanged from Go: 4 x u64 little-endian limbs, value * 2^256 mod p, canonical.
//
// The multiplier is the operand-scanning Montgomery product of csrc/fr_device.cuh (two accumulators split by the parity of the
// limb position, carry-chained IMAD.WIDE.U32) with the modulus as a template parameter.  On the device the carry chains ARE
// fr_device.cuh's inline-PTX primitives (chain4, chain4_cin, add8_carry, fr_sqr_wide), the ones every GKR parity test exercises;
// compiled for the host (`__CUDA_ARCH__` undefined) the same primitives are plain 64-bit C++ with identical semantics, which is
// what lets tests/test_msm_cpu.py run every kernel body of the multi-exponentiation on the CPU against the oracle.  The host
// form is also what the library itself uses for its few host-side field operations (RawBytes of the final point).
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#include "../fr_device.cuh"
#define EC_HD __host__ __device__ __forceinline__
#define EC_HD_NOINLINE __host__ __device__ __noinline__
#else
#define EC_HD inline
#define EC_HD_NOINLINE inline
#endif

namespace ec {

struct Big8 {
    uint32_t v[8];
};

// ---- carry-chain primitives -------------------------------------------------------------------------------------------------
// (c1:c0) += x0*y, (c3:c2) += x1*y, (c5:c4) += x2*y, (c7:c6) += x3*y as ONE 256-bit addition, carry out added to `top`.
EC_HD void chain4(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& c4, uint32_t& c5, uint32_t& c6, uint32_t& c7, uint32_t& top,
                  uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
#ifdef __CUDA_ARCH__
    gkr::chain4(c0, c1, c2, c3, c4, c5, c6, c7, top, x0, x1, x2, x3, y);
#else
    unsigned __int128 t;
    uint64_t carry;
    t = (unsigned __int128)(((uint64_t)c1 << 32) | c0) + (uint64_t)x0 * y;
    c0 = (uint32_t)t, c1 = (uint32_t)(t >> 32), carry = (uint64_t)(t >> 64);
    t = (unsigned __int128)(((uint64_t)c3 << 32) | c2) + (uint64_t)x1 * y + carry;
    c2 = (uint32_t)t, c3 = (uint32_t)(t >> 32), carry = (uint64_t)(t >> 64);
    t = (unsigned __int128)(((uint64_t)c5 << 32) | c4) + (uint64_t)x2 * y + carry;
    c4 = (uint32_t)t, c5 = (uint32_t)(t >> 32), carry = (uint64_t)(t >> 64);
    t = (unsigned __int128)(((uint64_t)c7 << 32) | c6) + (uint64_t)x3 * y + carry;
    c6 = (uint32_t)t, c7 = (uint32_t)(t >> 32), carry = (uint64_t)(t >> 64);
    top += (uint32_t)carry;
#endif
}
// the same with carry-in = carry of the 32-bit addition d0 + d1
EC_HD void chain4_cin(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& c4, uint32_t& c5, uint32_t& c6, uint32_t& c7, uint32_t& top,
                      uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y, uint32_t d0, uint32_t d1) {
#ifdef __CUDA_ARCH__
    gkr::chain4_cin(c0, c1, c2, c3, c4, c5, c6, c7, top, x0, x1, x2, x3, y, d0, d1);
#else
    unsigned __int128 t;
    uint64_t carry = ((uint64_t)d0 + d1) >> 32;
    t = (unsigned __int128)(((uint64_t)c1 << 32) | c0) + (uint64_t)x0 * y + carry;
    c0 = (uint32_t)t, c1 = (uint32_t)(t >> 32), carry = (uint64_t)(t >> 64);
    t = (unsigned __int128)(((uint64_t)c3 << 32) | c2) + (uint64_t)x1 * y + carry;
    c2 = (uint32_t)t, c3 = (uint32_t)(t >> 32), carry = (uint64_t)(t >> 64);
    t = (unsigned __int128)(((uint64_t)c5 << 32) | c4) + (uint64_t)x2 * y + carry;
    c4 = (uint32_t)t, c5 = (uint32_t)(t >> 32), carry = (uint64_t)(t >> 64);
    t = (unsigned __int128)(((uint64_t)c7 << 32) | c6) + (uint64_t)x3 * y + carry;
    c6 = (uint32_t)t, c7 = (uint32_t)(t >> 32), carry = (uint64_t)(t >> 64);
    top += (uint32_t)carry;
#endif
}
// a[0..8) += b[0..8) (+ cin), returns the carry out
EC_HD uint32_t add8_carry_in(uint32_t* a, const uint32_t* b, uint32_t cin) {
#ifdef __CUDA_ARCH__
    return gkr::add8_carry_in(a, b, cin);
#else
    uint64_t c = cin;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a[i] + b[i];
        a[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
#endif
}
EC_HD uint32_t add8_carry(uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
    return gkr::add8_carry(a, b);
#else
    return add8_carry_in(a, b, 0);
#endif
}
// w[0..16) = a^2 as a plain 512-bit integer (36 wide multiply-adds on the device: fr_device.cuh's dedicated squaring)
EC_HD void sqr_wide(uint32_t (&w)[16], const Big8& a) {
#ifdef __CUDA_ARCH__
    gkr::Fr x;
#pragma unroll
    for (int i = 0; i < 8; i++) x.v[i] = a.v[i];
    gkr::fr_sqr_wide(w, x);
#else
    uint64_t acc[17];
    for (int i = 0; i < 17; i++) acc[i] = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < 8; j++) {
            const unsigned __int128 t = (unsigned __int128)a.v[i] * a.v[j] + acc[i + j] + carry;
            acc[i + j] = (uint32_t)t;
            carry = (uint64_t)(t >> 32);
        }
        acc[i + 8] += carry;
    }
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)acc[i];
#endif
}

// ---- moduli -----------------------------------------------------------------------------------------------------------------
// BN254 base field p = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47 (gnark-crypto ecc/bn254/fp)
struct FpMod {
    static constexpr uint32_t M0 = 0xd87cfd47u, M1 = 0x3c208c16u, M2 = 0x6871ca8du, M3 = 0x97816a91u, M4 = 0x8181585du, M5 = 0xb85045b6u,
                              M6 = 0xe131a029u, M7 = 0x30644e72u;
    static constexpr uint32_t NINV = 0xe4866389u;  // -p^-1 mod 2^32
    // 2^256 mod p (Montgomery one) and 2^512 mod p
    static constexpr uint32_t R0 = 0xc58f0d9du, R1 = 0xd35d438du, R2 = 0xf5c70b3du, R3 = 0x0a78eb28u, R4 = 0x7879462cu, R5 = 0x666ea36fu,
                              R6 = 0x9a07df2fu, R7 = 0x0e0a77c1u;
    static constexpr uint32_t RR0 = 0x538afa89u, RR1 = 0xf32cfc5bu, RR2 = 0xd44501fbu, RR3 = 0xb5e71911u, RR4 = 0x0a417ff6u, RR5 = 0x47ab1effu,
                              RR6 = 0xcab8351fu, RR7 = 0x06d89f71u;
};
// BN254 scalar field q = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001 (ecc/bn254/fr)
struct FrMod {
    static constexpr uint32_t M0 = 0xf0000001u, M1 = 0x43e1f593u, M2 = 0x79b97091u, M3 = 0x2833e848u, M4 = 0x8181585du, M5 = 0xb85045b6u,
                              M6 = 0xe131a029u, M7 = 0x30644e72u;
    static constexpr uint32_t NINV = 0xefffffffu;
    static constexpr uint32_t R0 = 0x4ffffffbu, R1 = 0xac96341cu, R2 = 0x9f60cd29u, R3 = 0x36fc7695u, R4 = 0x7879462eu, R5 = 0x666ea36fu,
                              R6 = 0x9a07df2fu, R7 = 0x0e0a77c1u;
    static constexpr uint32_t RR0 = 0xae216da7u, RR1 = 0x1bb8e645u, RR2 = 0xe35c59e3u, RR3 = 0x53fe3ab1u, RR4 = 0x53bb8085u, RR5 = 0x8c49833du,
                              RR6 = 0x7f4e44a5u, RR7 = 0x0216d0b1u;
};

#define EC_MOD_LIMBS(F) {F::M0, F::M1, F::M2, F::M3, F::M4, F::M5, F::M6, F::M7}

EC_HD Big8 big_zero() {
    Big8 r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
EC_HD bool big_is_zero(const Big8& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7]) == 0; }
EC_HD bool big_eq(const Big8& a, const Big8& b) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d |= a.v[i] ^ b.v[i];
    return d == 0;
}
template <class F>
EC_HD Big8 f_one() {
    Big8 r;
    r.v[0] = F::R0, r.v[1] = F::R1, r.v[2] = F::R2, r.v[3] = F::R3, r.v[4] = F::R4, r.v[5] = F::R5, r.v[6] = F::R6, r.v[7] = F::R7;
    return r;
}
template <class F>
EC_HD Big8 f_rsquare() {
    Big8 r;
    r.v[0] = F::RR0, r.v[1] = F::RR1, r.v[2] = F::RR2, r.v[3] = F::RR3, r.v[4] = F::RR4, r.v[5] = F::RR5, r.v[6] = F::RR6, r.v[7] = F::RR7;
    return r;
}
// a - m with the borrow (all ones when a < m, else 0); plain C++ on both sides: additions and subtractions are a small share of
// the work next to the multiplier
template <class F>
EC_HD uint32_t f_sub_mod(Big8& r, const Big8& a) {
    const uint32_t m[8] = EC_MOD_LIMBS(F);
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint64_t d = (uint64_t)a.v[i] - m[i] - borrow;
        r.v[i] = (uint32_t)d;
        borrow = (d >> 32) & 1;
    }
    return 0u - (uint32_t)borrow;
}
// canonicalise a value known to be < 2m
template <class F>
EC_HD Big8 f_reduce_once(const Big8& a) {
    Big8 s;
    const uint32_t borrow = f_sub_mod<F>(s, a);
    Big8 r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = borrow ? a.v[i] : s.v[i];
    return r;
}
template <class F>
EC_HD bool f_is_canonical(const Big8& a) {
    Big8 s;
    return f_sub_mod<F>(s, a) != 0;
}
template <class F>
EC_HD Big8 f_add(const Big8& a, const Big8& b) {
    Big8 t;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] + b.v[i];
        t.v[i] = (uint32_t)c;
        c >>= 32;
    }
    return f_reduce_once<F>(t);  // a + b < 2m < 2^255
}
template <class F>
EC_HD Big8 f_sub(const Big8& a, const Big8& b) {
    const uint32_t m[8] = EC_MOD_LIMBS(F);
    Big8 t;
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint64_t d = (uint64_t)a.v[i] - b.v[i] - borrow;
        t.v[i] = (uint32_t)d;
        borrow = (d >> 32) & 1;
    }
    const uint32_t mask = 0u - (uint32_t)borrow;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)t.v[i] + (m[i] & mask);
        t.v[i] = (uint32_t)c;
        c >>= 32;
    }
    return t;
}
template <class F>
EC_HD Big8 f_dbl(const Big8& a) {
    return f_add<F>(a, a);
}
template <class F>
EC_HD Big8 f_neg(const Big8& a) {
    return f_sub<F>(big_zero(), a);
}

// Montgomery product a*b*2^-256 mod m; inputs canonical, output canonical.  Row structure of gkr::fr_mul_school (fr_device.cuh):
// P[p] / Qd[p] are the accumulator limbs at absolute position p of the accumulators whose 64-bit columns start at even / odd
// positions; row i adds a*b_i*2^(32i) and k_i*m*2^(32i), which zeroes limb i of P + Qd; its carry rides into the next chain.
template <class F>
EC_HD Big8 f_mul(const Big8& a, const Big8& b) {
    uint32_t P[18], Qd[18];
#pragma unroll
    for (int i = 0; i < 18; i++) P[i] = 0, Qd[i] = 0;
    const uint32_t q[8] = EC_MOD_LIMBS(F);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t* S = (i & 1) ? Qd : P;
        uint32_t* T = (i & 1) ? P : Qd;
        const uint32_t bi = b.v[i];
        chain4(S[i], S[i + 1], S[i + 2], S[i + 3], S[i + 4], S[i + 5], S[i + 6], S[i + 7], S[i + 8], a.v[0], a.v[2], a.v[4], a.v[6], bi);
        const uint32_t k = (S[i] + T[i]) * F::NINV;
        chain4(S[i], S[i + 1], S[i + 2], S[i + 3], S[i + 4], S[i + 5], S[i + 6], S[i + 7], S[i + 8], q[0], q[2], q[4], q[6], k);
        chain4_cin(T[i + 1], T[i + 2], T[i + 3], T[i + 4], T[i + 5], T[i + 6], T[i + 7], T[i + 8], T[i + 9], a.v[1], a.v[3], a.v[5], a.v[7], bi,
                   S[i], T[i]);
        chain4(T[i + 1], T[i + 2], T[i + 3], T[i + 4], T[i + 5], T[i + 6], T[i + 7], T[i + 8], T[i + 9], q[1], q[3], q[5], q[7], k);
    }
    Big8 t;
#pragma unroll
    for (int i = 0; i < 8; i++) t.v[i] = P[8 + i];
    (void)add8_carry(t.v, Qd + 8);  // (a*b + K*m) / 2^256 < 2m < 2^255: no carry out
    return f_reduce_once<F>(t);
}
// Montgomery reduction of a 512-bit t < m * 2^256 (gkr::fr_redc_wide with the modulus as a parameter)
template <class F>
EC_HD Big8 f_redc_wide(const uint32_t (&t)[16]) {
    uint32_t P[18], Qd[18];
#pragma unroll
    for (int i = 0; i < 18; i++) P[i] = i < 8 ? t[i] : 0, Qd[i] = 0;
    const uint32_t q[8] = EC_MOD_LIMBS(F);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t* S = (i & 1) ? Qd : P;
        uint32_t* T = (i & 1) ? P : Qd;
        const uint32_t k = (S[i] + T[i]) * F::NINV;
        chain4(S[i], S[i + 1], S[i + 2], S[i + 3], S[i + 4], S[i + 5], S[i + 6], S[i + 7], S[i + 8], q[0], q[2], q[4], q[6], k);
        chain4_cin(T[i + 1], T[i + 2], T[i + 3], T[i + 4], T[i + 5], T[i + 6], T[i + 7], T[i + 8], T[i + 9], q[1], q[3], q[5], q[7], k, S[i], T[i]);
    }
    Big8 r;
    uint32_t hi[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P[8 + i], hi[i] = t[8 + i];
    (void)add8_carry(r.v, Qd + 8);
    (void)add8_carry(r.v, hi);  // (t + K*m) / 2^256 < 2m: no carry out
    return f_reduce_once<F>(r);
}
// Square: 36 + 72 = 108 wide multiply-adds instead of 136
template <class F>
EC_HD Big8 f_sqr(const Big8& a) {
    uint32_t w[16];
    sqr_wide(w, a);
    return f_redc_wide<F>(w);
}
// out-of-line copies for the latency-bound single-thread kernels (small code)
template <class F>
EC_HD_NOINLINE Big8 f_mulc(const Big8 a, const Big8 b) {
    return f_mul<F>(a, b);
}
template <class F>
EC_HD Big8 f_from_mont(const Big8& a) {
    Big8 one = big_zero();
    one.v[0] = 1;
    return f_mul<F>(a, one);
}
template <class F>
EC_HD Big8 f_to_mont(const Big8& a) {
    return f_mul<F>(a, f_rsquare<F>());
}
// a^-1 = a^(m-2) (0 -> 0), square-and-multiply from the top bit; one call per multi-exponentiation (the final affine conversion)
template <class F>
EC_HD_NOINLINE Big8 f_inv(const Big8 a) {
    uint32_t e[8] = EC_MOD_LIMBS(F);
    e[0] -= 2;  // both moduli end in ...01 / ...47: no borrow
    Big8 acc = f_one<F>();
    for (int i = 253; i >= 0; i--) {  // m < 2^254
        acc = f_mulc<F>(acc, acc);
        if ((e[i >> 5] >> (i & 31)) & 1) acc = f_mulc<F>(acc, a);
    }
    return acc;
}

// memory images: 4 x u64 little-endian == 8 x u32 little-endian
EC_HD Big8 big_load(const uint64_t* p) {
    Big8 r;
#if defined(__CUDA_ARCH__)
    const uint4 lo = reinterpret_cast<const uint4*>(p)[0], hi = reinterpret_cast<const uint4*>(p)[1];
    r.v[0] = lo.x, r.v[1] = lo.y, r.v[2] = lo.z, r.v[3] = lo.w, r.v[4] = hi.x, r.v[5] = hi.y, r.v[6] = hi.z, r.v[7] = hi.w;
#else
    for (int i = 0; i < 4; i++) r.v[2 * i] = (uint32_t)p[i], r.v[2 * i + 1] = (uint32_t)(p[i] >> 32);
#endif
    return r;
}
EC_HD void big_store(uint64_t* p, const Big8& a) {
#if defined(__CUDA_ARCH__)
    reinterpret_cast<uint4*>(p)[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    reinterpret_cast<uint4*>(p)[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
#else
    for (int i = 0; i < 4; i++) p[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
#endif
}

}  // namespace ec
