// Host-side serial pieces of the prover: MiMC Fiat-Shamir hash, Lagrange interpolation, eq closed form.
// These stay on the CPU on purpose: each sumcheck round needs 9 blocks x 91 rounds x 4 = 3276 strictly
// dependent field multiplications (hash/mimc.go:11-49); one x86 core runs that chain several times
// faster than one GPU thread can (see DESIGN.md "transcript").
#pragma once
#include <vector>

#include "fr_host.hpp"

namespace gkr {
namespace host {

static constexpr int MIMC_ROUNDS = 91;  // hash/mimc.go:8

static const Fr ARKS[MIMC_ROUNDS] = {
#include "mimc_arks.inc"
};

// hash/mimc.go:31-39 MimcKeyedPermutation.  x^7 is evaluated as x^4 * x^3 with x^3 and x^4 independent
// (dependency depth 3 instead of the reference's 4-long chain x^2,x^3,x^6,x^7); exact arithmetic, same value.
// Inside a round the products stay unreduced in [0, 2q): with res < q and key+ark < q,
//   t = res + (key+ark) < 2q,  t^2 < 1.76q,  t^4 < 1.59q,  t^3 < 1.67q,  t^4*t^3 < 1.50q  (u*v/2^256 + q, q/2^256 = 0.189),
// so ONE conditional subtraction per round restores res < q; the value mod q is what the reference computes.
static inline Fr mimc_keyed_permutation(const Fr& x, const Fr& key) {
    Fr ka[MIMC_ROUNDS];  // key + ark_i: off the multiplication chain
    for (int i = 0; i < MIMC_ROUNDS; i++) ka[i] = add(key, ARKS[i]);
    Fr res = x;
    for (int i = 0; i < MIMC_ROUNDS; i++) {
        const Fr t = add_lazy(res, ka[i]);
        const Fr t2 = mul_lazy(t, t);
        const Fr t4 = mul_lazy(t2, t2);
        const Fr t3 = mul_lazy(t2, t);
        const Fr r = mul_lazy(t4, t3);
        res = reduce_once(r.l[0], r.l[1], r.l[2], r.l[3]);
    }
    return res;
}
// hash/mimc.go:24-28 MimcUpdateInplace (Miyaguchi-Preneel): s <- s + (Perm_s(b) + s) + b
static inline void mimc_update(Fr& state, const Fr& block) {
    Fr ns = add(mimc_keyed_permutation(block, state), state);
    state = add(add(state, ns), block);
}
// hash/mimc.go:11-18 MimcHash == common/challenge.go:10 GetChallenge
static inline Fr mimc_hash(const Fr* in, size_t n) {
    Fr s = zero();
    for (size_t i = 0; i < n; i++) mimc_update(s, in[i]);
    return s;
}

// poly/lagrange.go:31-39 EvalUnivariate (coefficients low -> high)
static inline Fr eval_univariate(const Fr* coeffs, size_t n, const Fr& x) {
    Fr res = coeffs[n - 1];
    for (size_t i = n - 1; i-- > 0;) res = add(mul(res, x), coeffs[i]);
    return res;
}

// poly/eq.go:19-32 EvalEq
static inline Fr eval_eq(const Fr* q, const Fr* h, size_t n) {
    Fr res = one();
    const Fr o = one();
    for (size_t i = 0; i < n; i++) {
        Fr nxt = mul(q[i], h[i]);
        nxt = add(add(nxt, nxt), o);
        nxt = sub(nxt, add(q[i], h[i]));
        res = mul(res, nxt);
    }
    return res;
}

// Coefficient form of the Lagrange basis on {0..d-1}: basis[i][j] = coeff of X^j in L_i(X).
// Same matrices as poly/lagrange.go:42-92 (they are unique); built as prod_{k!=i}(X-k) / prod_{k!=i}(i-k).
struct Lagrange {
    static constexpr int MAX_DOMAIN = 12;  // poly/lagrange.go:21
    Fr basis[MAX_DOMAIN + 1][MAX_DOMAIN][MAX_DOMAIN];
    Lagrange() {
        for (int d = 1; d <= MAX_DOMAIN; d++) {
            for (int i = 0; i < d; i++) {
                Fr num[MAX_DOMAIN + 1];
                for (auto& c : num) c = zero();
                num[0] = one();
                int deg = 0;
                Fr den = one();
                for (int k = 0; k < d; k++) {
                    if (k == i) continue;
                    Fr mk = neg(from_u64((uint64_t)k));
                    // num *= (X - k)
                    for (int j = deg + 1; j >= 1; j--) num[j] = add(num[j - 1], mul(num[j], mk));
                    num[0] = mul(num[0], mk);
                    deg++;
                    Fr diff = i > k ? from_u64((uint64_t)(i - k)) : neg(from_u64((uint64_t)(k - i)));
                    den = mul(den, diff);
                }
                Fr dinv = inv(den);
                for (int j = 0; j < d; j++) basis[d][i][j] = mul(num[j], dinv);
            }
        }
    }
    // poly/lagrange.go:96-111 InterpolateOnRange: evaluations on 0..n-1 -> coefficients low->high
    void interpolate(const Fr* values, int n, Fr* out) const {
        for (int j = 0; j < n; j++) out[j] = zero();
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) out[j] = add(out[j], mul(basis[n][i][j], values[i]));
    }
};

}  // namespace host
}  // namespace gkr
