// fft.Domain.FFT / FFTInverse and computeH on the device: radix-2 number-theoretic transforms over BN254 Fr, three stages per pass.
//
// Replaces gnark-crypto's ecc/bn254/fr/fft (reference go.mod:7, un-vendored: recursive difFFT / ditFFT, one goroutine per split)
// behind computeH (prover/gadget/prove.go:310-366: 3 x FFTInverse(DIF), 3 x FFT(DIT, coset 1), h = (a o b - c) / -2,
// FFTInverse(DIF, coset 1), FromMont) on the domain of fft.NewDomain(len(r1cs.Constraints), 1, true)
// (pkg/gnark/notinternal/backend/bn254/groth16/setup.go:98).
//
// Layout: a transform works IN PLACE on n contiguous fr.Element images (32 B, Montgomery), exactly Go's []fr.Element.  As in the
// reference no bit-reversal pass ever runs: DIF leaves its output bit-reversed, DIT takes bit-reversed input (computeH chains them).
//   KNtt<R, DIF>  one thread owns 2^R elements  base + i * s  and runs R consecutive butterfly stages on them in registers
//                 (R = 3: 8 elements = 64 registers, 12 butterflies, ONE read and ONE write of the array for three stages: 8 passes
//                 for n = 2^22 instead of 22).  Neighbouring threads own neighbouring k, so every load / store of a warp is
//                 one contiguous 1 KiB run (or, at strides below 32 elements, whole 32-byte sectors).
//   twiddles      one table w^j, j < n/2, per direction (built once per domain by KTwiddle from two sqrt(n)-sized host tables);
//                 a stage of half-length m reads w_2m^j = tw[j * n/2m] -- the late DIF / early DIT stages touch few distinct
//                 entries, which stay in L1/L2.
//   KScale        a[i] *= g^e(i) * c, e(i) = i or rev(i): the coset tables of the reference (CosetTable / CosetTableReversed,
//                 n entries each there) are replaced by g^e = lo[e & 2047] * hi[e >> 11], 1/n folded into `hi`
//   KPointwise    a = (a o b - c) * k
// Field arithmetic is exact, so whatever the order of the butterflies every output is the same canonical residue as the reference's.
// Every body is a function of its thread index (no shared memory, no cooperation): tests/emu/msm_emu.cpp runs the same bodies and the
// same drivers below on the CPU against the oracle.
#pragma once
#include <vector>

#include "field.cuh"

namespace ec {

constexpr int NTT_LO_BITS = 11;  // g^e = lo[e & 2047] * hi[e >> 11]
constexpr int NTT_MAX_LOG = 26;  // 2^26 elements = 2 GiB per array; BN254 Fr has 2-adicity 28 and the coset needs one more bit

// 2^28-th primitive root of unity of BN254 Fr used by gnark-crypto's fft.NewDomain (= 5^((q-1)/2^28)), regular form:
// 19103219067921713944291392827692070036145651957329286315305642004821462161904
EC_HD Big8 ntt_root_of_unity_regular() {
    Big8 r;
    r.v[0] = 0x725b19f0u, r.v[1] = 0x9bd61b6eu, r.v[2] = 0x41112ed4u, r.v[3] = 0x402d111eu;
    r.v[4] = 0x8ef62abcu, r.v[5] = 0x00e0a7ebu, r.v[6] = 0xa58a7e85u, r.v[7] = 0x2a3c09f0u;
    return r;
}

EC_HD uint32_t ntt_rev(uint32_t x, uint32_t log_n) {  // bit reversal of the low log_n bits (log_n >= 1)
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    x = (x >> 16) | (x << 16);
    return x >> (32 - log_n);
}

// ---- kernel bodies ------------------------------------------------------------------------------------------------------------
struct KTwiddle {  // j < n/2: tw[j] = lo[j & 2047] * hi[j >> 11]
    static EC_HD void run(size_t j, uint64_t* tw, const uint64_t* lo, const uint64_t* hi) {
        const Big8 a = big_load(lo + 4 * (j & ((1u << NTT_LO_BITS) - 1))), b = big_load(hi + 4 * (j >> NTT_LO_BITS));
        big_store(tw + 4 * j, f_mul<FrMod>(a, b));
    }
};

// R butterfly stages on the 2^R elements base + i * s, s = 2^log_s, of every block of 2^(log_s + R) elements.
//   DIF (Gentleman-Sande): half-lengths 2^(R-1) s, ..., s in this order:  (u, v) -> (u + v, (u - v) w)
//   DIT (Cooley-Tukey):    half-lengths s, 2 s, ..., 2^(R-1) s:           (u, v) -> (u + v w, u - v w)
// w = w_2m^j = tw[j << (log_n - 1 - log m)], j = position of the pair inside its block of 2m.
template <int R, bool DIF>
struct KNtt {  // t < n >> R
    static EC_HD void run(size_t t, uint64_t* a, const uint64_t* tw, uint32_t log_n, uint32_t log_s) {
        constexpr int E = 1 << R;
        const size_t s = (size_t)1 << log_s;
        const size_t k = t & (s - 1);
        const size_t base = ((t >> log_s) << (log_s + R)) + k;
        Big8 x[E];
#pragma unroll
        for (int i = 0; i < E; i++) x[i] = big_load(a + 4 * (base + (size_t)i * s));
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int lg_half = DIF ? R - 1 - r : r;  // half-length of this stage in units of s
            const int half = 1 << lg_half;
            const uint32_t sh = log_n - 1 - (log_s + (uint32_t)lg_half);
#pragma unroll
            for (int i = 0; i < E; i++) {
                if (i & half) continue;
                const size_t j = (size_t)(i & (half - 1)) * s + k;
                const Big8 w = big_load(tw + 4 * (j << sh));
                if (DIF) {
                    const Big8 u = x[i], v = x[i + half];
                    x[i] = f_add<FrMod>(u, v);
                    x[i + half] = f_mul<FrMod>(f_sub<FrMod>(u, v), w);
                } else {
                    const Big8 u = x[i], v = f_mul<FrMod>(x[i + half], w);
                    x[i] = f_add<FrMod>(u, v);
                    x[i + half] = f_sub<FrMod>(u, v);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < E; i++) big_store(a + 4 * (base + (size_t)i * s), x[i]);
    }
};

struct KScale {  // i < n: a[i] *= lo[e & 2047] * hi[e >> 11], e = i or rev(i); optionally FromMont afterwards
    static EC_HD void run(size_t i, uint64_t* a, const uint64_t* lo, const uint64_t* hi, uint32_t log_n, uint32_t reversed, uint32_t from_mont) {
        const uint32_t e = (reversed && log_n) ? ntt_rev((uint32_t)i, log_n) : (uint32_t)i;
        const Big8 g = f_mul<FrMod>(big_load(lo + 4 * (e & ((1u << NTT_LO_BITS) - 1))), big_load(hi + 4 * (e >> NTT_LO_BITS)));
        Big8 v = f_mul<FrMod>(big_load(a + 4 * i), g);
        if (from_mont) v = f_from_mont<FrMod>(v);
        big_store(a + 4 * i, v);
    }
};
struct KMulConst {  // i < n: a[i] *= k[0]
    static EC_HD void run(size_t i, uint64_t* a, const uint64_t* k) { big_store(a + 4 * i, f_mul<FrMod>(big_load(a + 4 * i), big_load(k))); }
};
struct KPointwise {  // i < n: a[i] = (a[i] * b[i] - c[i]) * k[0]      (prove.go:346-352)
    static EC_HD void run(size_t i, uint64_t* a, const uint64_t* b, const uint64_t* c, const uint64_t* k) {
        const Big8 p = f_mul<FrMod>(big_load(a + 4 * i), big_load(b + 4 * i));
        big_store(a + 4 * i, f_mul<FrMod>(f_sub<FrMod>(p, big_load(c + 4 * i)), big_load(k)));
    }
};

// ---- domain (host side) -------------------------------------------------------------------------------------------------------
// The small tables of fft.NewDomain(m, 1, _) in Montgomery form, computed on the host with field.cuh's host multiplier:
// everything the device needs to build its twiddle tables and to scale by coset powers.
struct NttDomainHost {
    uint32_t log_n = 0;
    std::vector<uint64_t> w_lo, w_hi, wi_lo, wi_hi;  // Generator^e / GeneratorInv^e split (e < n/2)
    std::vector<uint64_t> u_lo, u_hi, u_hi_n;        // FinerGenerator^e split (e < n); u_hi_n carries the factor 1/n
    std::vector<uint64_t> ui_lo, ui_hi_n;            // FinerGeneratorInv^e / n
    uint64_t n_inv[4], minus_two_inv[4], generator[4], finer_generator[4];
};
inline Big8 ntt_pow2k(Big8 x, uint32_t k) {
    for (uint32_t i = 0; i < k; i++) x = f_sqr<FrMod>(x);
    return x;
}
inline void ntt_split_tables(const Big8& g, size_t count, const Big8& hi_factor, std::vector<uint64_t>& lo, std::vector<uint64_t>& hi) {
    const size_t nlo = (size_t)1 << NTT_LO_BITS;
    const size_t nhi = (count + nlo - 1) >> NTT_LO_BITS;
    lo.assign(4 * nlo, 0);
    hi.assign(4 * (nhi ? nhi : 1), 0);
    Big8 t = f_one<FrMod>();
    for (size_t i = 0; i < nlo; i++) {
        big_store(lo.data() + 4 * i, t);
        t = f_mul<FrMod>(t, g);
    }
    const Big8 step = t;  // g^2048
    t = hi_factor;
    for (size_t i = 0; i < (nhi ? nhi : 1); i++) {
        big_store(hi.data() + 4 * i, t);
        t = f_mul<FrMod>(t, step);
    }
}
inline bool ntt_domain_host(uint32_t log_n, NttDomainHost& d) {
    if (log_n > (uint32_t)NTT_MAX_LOG) return false;
    d.log_n = log_n;
    const size_t n = (size_t)1 << log_n;
    const Big8 g = f_to_mont<FrMod>(ntt_root_of_unity_regular());
    const Big8 w = ntt_pow2k(g, 28 - log_n), u = ntt_pow2k(g, 28 - log_n - 1);  // Generator, FinerGenerator (depth 1): u^2 = w
    const Big8 wi = f_inv<FrMod>(w), ui = f_inv<FrMod>(u);
    Big8 nn = big_zero();
    nn.v[0] = (uint32_t)n;  // n <= 2^26
    const Big8 n_inv = f_inv<FrMod>(f_to_mont<FrMod>(nn));
    const Big8 one = f_one<FrMod>();
    const size_t half = n > 1 ? n / 2 : 1;
    ntt_split_tables(w, half, one, d.w_lo, d.w_hi);
    ntt_split_tables(wi, half, one, d.wi_lo, d.wi_hi);
    ntt_split_tables(u, n, one, d.u_lo, d.u_hi);
    std::vector<uint64_t> dummy;
    ntt_split_tables(u, n, n_inv, dummy, d.u_hi_n);
    ntt_split_tables(ui, n, n_inv, d.ui_lo, d.ui_hi_n);
    Big8 two = big_zero();
    two.v[0] = 2;
    const Big8 m2i = f_inv<FrMod>(f_neg<FrMod>(f_to_mont<FrMod>(two)));  // (-2)^-1   (prove.go:340-343)
    big_store(d.n_inv, n_inv);
    big_store(d.minus_two_inv, m2i);
    big_store(d.generator, w);
    big_store(d.finer_generator, u);
    return true;
}

// device-resident domain: pointers into one buffer (see ntt_domain_layout)
struct NttDomainDev {
    uint32_t log_n;
    uint64_t *tw, *tw_inv;                                  // n/2 entries each
    uint64_t *w_lo, *w_hi, *wi_lo, *wi_hi;                   // only needed to build tw / tw_inv
    uint64_t *u_lo, *u_hi, *u_hi_n, *ui_lo, *ui_hi_n;
    uint64_t *n_inv, *minus_two_inv;
};

// ---- drivers ------------------------------------------------------------------------------------------------------------------
// in-place transform of the n = 2^log_n elements at `a` with the twiddle table `tw` (forward or inverse powers); returns launches
template <class Exec>
int ntt_enqueue(Exec& ex, uint64_t* a, const uint64_t* tw, uint32_t log_n, bool dif) {
    int launches = 0;
    const size_t n = (size_t)1 << log_n;
    if (dif) {
        uint32_t left = log_n;  // stages still to run; the next one has half-length 2^(left-1)
        while (left) {
            const uint32_t R = left >= 3 ? 3 : left, log_s = left - R;
            if (R == 3) launches += ex.template launch<KNtt<3, true>>(n >> 3, a, tw, log_n, log_s);
            else if (R == 2) launches += ex.template launch<KNtt<2, true>>(n >> 2, a, tw, log_n, log_s);
            else launches += ex.template launch<KNtt<1, true>>(n >> 1, a, tw, log_n, log_s);
            left -= R;
        }
    } else {
        uint32_t done = 0;  // stages run so far; the next one has half-length 2^done
        while (done < log_n) {
            const uint32_t R = log_n - done >= 3 ? 3 : log_n - done, log_s = done;
            if (R == 3) launches += ex.template launch<KNtt<3, false>>(n >> 3, a, tw, log_n, log_s);
            else if (R == 2) launches += ex.template launch<KNtt<2, false>>(n >> 2, a, tw, log_n, log_s);
            else launches += ex.template launch<KNtt<1, false>>(n >> 1, a, tw, log_n, log_s);
            done += R;
        }
    }
    return launches;
}
// fft.Domain.FFT(a, decimation, coset) / FFTInverse, coset in {0, 1}
template <class Exec>
int fft_enqueue(Exec& ex, const NttDomainDev& d, uint64_t* a, bool dif, int coset, bool inverse) {
    int launches = 0;
    const size_t n = (size_t)1 << d.log_n;
    if (!inverse) {
        // the coset shift multiplies coefficient j by u^j before the transform; a DIT input holds coefficient rev(i) at i
        if (coset) launches += ex.template launch<KScale>(n, a, (const uint64_t*)d.u_lo, (const uint64_t*)d.u_hi, d.log_n, dif ? 0u : 1u, 0u);
        launches += ntt_enqueue(ex, a, (const uint64_t*)d.tw, d.log_n, dif);
    } else {
        launches += ntt_enqueue(ex, a, (const uint64_t*)d.tw_inv, d.log_n, dif);
        // 1/n, and u^-j on coefficient j; a DIF output holds coefficient rev(i) at i
        if (coset) launches += ex.template launch<KScale>(n, a, (const uint64_t*)d.ui_lo, (const uint64_t*)d.ui_hi_n, d.log_n, dif ? 1u : 0u, 0u);
        else launches += ex.template launch<KMulConst>(n, a, (const uint64_t*)d.n_inv);
    }
    return launches;
}
// computeH (prove.go:310-366) on three zero-padded arrays of n elements; the result replaces `a`: regular form, bit-reversed
// coefficient order (what the reference hands to MultiExp against the bit-reversed pk.G1.Z, setup.go:229).
// The 1/n of FFTInverse(DIF, 0) and the u^rev(i) of FFT(DIT, 1) are applied as ONE scaling (multiplication in Fr is exact and
// commutative: same residues), and FromMont rides on the last scaling pass.
template <class Exec>
int compute_h_enqueue(Exec& ex, const NttDomainDev& d, uint64_t* a, uint64_t* b, uint64_t* c) {
    int launches = 0;
    const size_t n = (size_t)1 << d.log_n;
    uint64_t* v[3] = {a, b, c};
    for (int t = 0; t < 3; t++) {
        launches += ntt_enqueue(ex, v[t], (const uint64_t*)d.tw_inv, d.log_n, true);                                        // FFTInverse(DIF, 0) ...
        launches += ex.template launch<KScale>(n, v[t], (const uint64_t*)d.u_lo, (const uint64_t*)d.u_hi_n, d.log_n, 1u, 0u);  // ... / n, coset 1
        launches += ntt_enqueue(ex, v[t], (const uint64_t*)d.tw, d.log_n, false);                                           // FFT(DIT, 1)
    }
    launches += ex.template launch<KPointwise>(n, a, (const uint64_t*)b, (const uint64_t*)c, (const uint64_t*)d.minus_two_inv);
    launches += ntt_enqueue(ex, a, (const uint64_t*)d.tw_inv, d.log_n, true);                                               // FFTInverse(DIF, 1)
    launches += ex.template launch<KScale>(n, a, (const uint64_t*)d.ui_lo, (const uint64_t*)d.ui_hi_n, d.log_n, 1u, 1u);       // + FromMont
    return launches;
}
// builds tw and tw_inv from the uploaded split tables
template <class Exec>
int ntt_domain_enqueue(Exec& ex, const NttDomainDev& d) {
    const size_t half = ((size_t)1 << d.log_n) / 2;
    int launches = 0;
    launches += ex.template launch<KTwiddle>(half, d.tw, (const uint64_t*)d.w_lo, (const uint64_t*)d.w_hi);
    launches += ex.template launch<KTwiddle>(half, d.tw_inv, (const uint64_t*)d.wi_lo, (const uint64_t*)d.wi_hi);
    return launches;
}

}  // namespace ec
