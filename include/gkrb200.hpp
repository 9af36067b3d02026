// gkrb200.hpp -- C++ mirror of the reference's Go API for the prover hot path, over the C ABI of gkrb200.h.
//
// The reference is Go and there is no Go toolchain in this image, so the host side above the C ABI is C++ (plus the Python
// mirror the tests use).  Package -> namespace, same names and argument meaning, so a test written against this header reads
// like the reference's own (`tests/cpp/test_api.cpp` restates gkr/gkr_test.go, sumcheck/prover_test.go, poly/multilin_test.go
// and hash/hash_test.go).  Where Go panics (sumcheck/prover.go:54,114; gkr/prover.go:84; poly/pool.go:71) this header throws
// gkrmimc::Panic carrying gkrb200_last_error(); where Go returns `error` (gkr.Verify) it returns a non-empty string.
//
//   common.GetChallenge / RandomFrArray      common/challenge.go:10, common/common.go:49-55
//   hash.MimcHash                            hash/mimc.go:11-18
//   poly.MultiLin / Fold / Evaluate / FoldedEqTable / InterpolateOnRange
//                                            poly/multilin.go:19-66, poly/eq.go:41-59, poly/lagrange.go:96-111
//   circuit.Gate / Layer / Circuit / BuildCircuit / Assign / Assignment
//                                            circuit/gates.go:9-21, circuit/circuit.go:12-91, circuit/assignment.go:12-57
//   gates.IdentityGate / NewCipherGate       circuit/gates/copy.go:9-32, circuit/gates/cipher.go:11-70
//   examples.MimcCircuit                     examples/mimc.go:10-37
//   sumcheck.Prove                           sumcheck/prover.go:46-90
//   gkr.Proof / Prove / Verify               gkr/prover.go:14-47, gkr/verifier.go:15-59
//   gadget.GkrProofToVec                     prover/gadget/hints.go:236-271
//
// Header-only; link with libgkrb200.so.  No CPU fallback: without a CUDA device Device() throws.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "gkrb200.h"

namespace gkrmimc {

struct Panic : std::runtime_error {
    int code;
    Panic(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
    if (rc != 0) throw Panic(rc, gkrb200_last_error());
}

namespace fr {
// fr.Element: [4]uint64 little-endian limbs, Montgomery form, canonical (gnark-crypto ecc/bn254/fr)
using Element = std::array<uint64_t, 4>;
inline Element SetUint64(uint64_t v) {
    Element in{v, 0, 0, 0}, out;
    check(gkrb200_to_montgomery(in.data(), 1, out.data()));
    return out;
}
inline Element ToRegular(const Element& x) {
    Element out;
    check(gkrb200_from_montgomery(x.data(), 1, out.data()));
    return out;
}
// scalar fr.Element methods (host): Mul, Add, Sub, the S-box x^7, Inverse
inline Element scalar(int op, const Element& a, const Element& b) {
    Element out;
    check(gkrb200_fr_scalar(op, a.data(), b.data(), out.data()));
    return out;
}
inline Element Mul(const Element& a, const Element& b) { return scalar(0, a, b); }
inline Element Add(const Element& a, const Element& b) { return scalar(1, a, b); }
inline Element Sub(const Element& a, const Element& b) { return scalar(2, a, b); }
inline Element Exp7(const Element& a) { return scalar(3, a, a); }
inline Element Inverse(const Element& a) { return scalar(4, a, a); }
inline Element One() { return SetUint64(1); }
}  // namespace fr

namespace detail {
// pointer to the elements of a (possibly empty) vector: an empty Go slice still has a valid base pointer, so the ABI never sees NULL for one
template <class V>
inline auto ptr(V& v) -> decltype(v[0].data()) {
    static fr::Element scratch{};
    return v.empty() ? const_cast<decltype(v[0].data())>(scratch.data()) : v[0].data();
}
}  // namespace detail

// One device context = the reference's package-global pool + worker goroutines (poly/pool.go, sumcheck/worker.go).
class Device {
public:
    explicit Device(int device = 0, int max_bn = 16) { check(gkrb200_init(&ctx_, device, max_bn, nullptr)); }
    ~Device() { gkrb200_free(ctx_); }
    Device(const Device&) = delete;
    Device& operator=(const Device&) = delete;
    gkrb200_ctx* handle() const { return ctx_; }

private:
    gkrb200_ctx* ctx_ = nullptr;
};

namespace common {
// common.RandomFrArray: SetUint64((i*i) ^ 0xf45c9df123f)
inline std::vector<fr::Element> RandomFrArray(size_t n) {
    std::vector<fr::Element> in(n), out(n);
    for (size_t i = 0; i < n; i++) in[i] = {((uint64_t)i * (uint64_t)i) ^ 0xf45c9df123fULL, 0, 0, 0};
    if (n) check(gkrb200_to_montgomery(in[0].data(), n, out[0].data()));
    return out;
}
// common.GetChallenge(seed) == hash.MimcHash(seed)
inline fr::Element GetChallenge(const std::vector<fr::Element>& seed) {
    fr::Element out;
    check(gkrb200_mimc_hash(detail::ptr(seed), seed.size(), out.data()));
    return out;
}
}  // namespace common

namespace hash {
inline fr::Element MimcHash(const std::vector<fr::Element>& in) { return common::GetChallenge(in); }
}  // namespace hash

namespace poly {
using MultiLin = std::vector<fr::Element>;
// MultiLin.Fold(r): the table keeps its first half (poly/multilin.go:19-23)
inline void Fold(Device& d, MultiLin& m, const fr::Element& r) {
    MultiLin out(m.size() / 2);
    check(gkrb200_fold(d.handle(), m[0].data(), m.size(), r.data(), detail::ptr(out)));
    m.swap(out);
}
inline fr::Element Evaluate(Device& d, const MultiLin& m, const std::vector<fr::Element>& coordinates) {
    fr::Element out;
    check(gkrb200_mle_evaluate(d.handle(), m[0].data(), m.size(), detail::ptr(coordinates), out.data()));
    return out;
}
inline MultiLin FoldedEqTable(Device& d, const std::vector<fr::Element>& qPrime) {
    MultiLin out((size_t)1 << qPrime.size());
    check(gkrb200_eq_table(d.handle(), detail::ptr(qPrime), 1, (int)qPrime.size(), nullptr, out[0].data()));
    return out;
}
inline std::vector<fr::Element> InterpolateOnRange(const std::vector<fr::Element>& values) {
    std::vector<fr::Element> out(values.size());
    check(gkrb200_interpolate(detail::ptr(values), values.size(), detail::ptr(out)));
    return out;
}
// poly.EvalUnivariate (poly/lagrange.go:31-39), coefficients low -> high
inline fr::Element EvalUnivariate(const std::vector<fr::Element>& coeffs, const fr::Element& x) {
    fr::Element out;
    check(gkrb200_eval_univariate(detail::ptr(coeffs), coeffs.size(), x.data(), out.data()));
    return out;
}
// poly.EvalEq (poly/eq.go:19-32)
inline fr::Element EvalEq(const std::vector<fr::Element>& qPrime, const std::vector<fr::Element>& nextQPrime) {
    if (qPrime.size() != nextQPrime.size()) throw Panic(GKRB200_ERR_ARG, "EvalEq: the two points have different sizes");
    fr::Element out;
    check(gkrb200_eval_eq(detail::ptr(qPrime), detail::ptr(nextQPrime), qPrime.size(), out.data()));
    return out;
}
}  // namespace poly

namespace circuit {
// circuit.Gate as it crosses the ABI: (kind, ark).  Arbitrary user gates cannot be expressed.
struct Gate {
    int kind = -1;  // -1: none (input layer), GKRB200_GATE_IDENTITY, GKRB200_GATE_CIPHER
    fr::Element ark{};
    std::string ID() const { return kind == GKRB200_GATE_CIPHER ? "cipher" : (kind == GKRB200_GATE_IDENTITY ? "identity" : ""); }
    int Degree() const { return kind == GKRB200_GATE_CIPHER ? 7 : 1; }  // cipher.go:68-70, copy.go:30-32
    bool nil() const { return kind < 0; }
    // Gate.Eval (cipher.go:45-55: (vL + vR + Ark)^7; copy.go:20-22: vL)
    fr::Element Eval(const std::vector<fr::Element>& xs) const {
        if (kind == GKRB200_GATE_IDENTITY) return xs.at(0);
        if (kind != GKRB200_GATE_CIPHER) throw Panic(GKRB200_ERR_ARG, "Eval on a nil gate");
        return fr::Exp7(fr::Add(fr::Add(xs.at(1), ark), xs.at(0)));
    }
};
struct Layer {
    Gate gate;
    std::vector<int> In, Out;
};
struct Circuit : std::vector<Layer> {
    using std::vector<Layer>::vector;  // make(circuit.Circuit, n)
    // circuit.go:70-79 (an input layer has no inputs and no gate; anything else inconsistent panics)
    bool IsInputLayer(int layer) const {
        const Layer& l = (*this)[(size_t)layer];
        if (l.In.empty() != l.gate.nil()) throw Panic(GKRB200_ERR_ARG, "layer is inconsistent : it should have either no gate and no inputs or a gate and inputs");
        return l.In.empty();
    }
    int InputArity() const {  // circuit.go:82-91
        int n = 0;
        for (size_t l = 0; l < size(); l++) n += IsInputLayer((int)l) ? 1 : 0;
        return n;
    }
};
// circuit.go:28-44: fills the Out lists; an input layer may feed a single layer
inline Circuit BuildCircuit(Circuit c) {
    for (size_t l = 0; l < c.size(); l++)
        for (int in : c[l].In) c[(size_t)in].Out.push_back((int)l);
    for (size_t l = 0; l < c.size(); l++)
        if (c.IsInputLayer((int)l) && c[l].Out.size() > 1) throw Panic(GKRB200_ERR_ARG, "input layer used by more than one layer");
    return c;
}

// circuit.Assignment: the 94 layer tables stay on the device; operator[] copies one to the host (what Go code reads as a[l])
class Assignment {
public:
    Assignment(Device& d, size_t n, int bn) : d_(&d), n_(n), bn_(bn) {}
    poly::MultiLin operator[](int layer) const {
        poly::MultiLin out(n_);
        check(gkrb200_assign_layer_to_host(d_->handle(), layer, out[0].data(), n_));
        return out;
    }
    // a[layer].Evaluate(coordinates) without copying the table off the device
    fr::Element Evaluate(int layer, const std::vector<fr::Element>& coordinates) const {
        fr::Element out;
        check(gkrb200_assign_layer_evaluate(d_->handle(), layer, detail::ptr(coordinates), bn_, out.data()));
        return out;
    }
    // Assignment.InputsOfLayer (circuit/assignment.go:46-57): host copies of the input tables of a layer
    template <class CircuitT>
    std::vector<poly::MultiLin> InputsOfLayer(const CircuitT& c, int layer) const {
        std::vector<poly::MultiLin> xs;
        for (int in : c[(size_t)layer].In) xs.push_back((*this)[in]);
        return xs;
    }
    size_t size() const { return GKRB200_MIMC_LAYERS; }
    int bn() const { return bn_; }
    Device& device() const { return *d_; }

private:
    Device* d_;
    size_t n_;
    int bn_;
};
}  // namespace circuit

namespace gates {
inline circuit::Gate IdentityGate() { return circuit::Gate{GKRB200_GATE_IDENTITY, {}}; }
inline circuit::Gate NewCipherGate(const fr::Element& ark) { return circuit::Gate{GKRB200_GATE_CIPHER, ark}; }
}  // namespace gates

namespace examples {
// hash.Arks[i] (hash/ark.go:13-337)
inline fr::Element Ark(int i) {
    fr::Element a;
    check(gkrb200_mimc_ark(i, a.data()));
    return a;
}
// examples/mimc.go:10-37.  Assign/Prove accept exactly this circuit (same wiring, same gates, same constants).
inline circuit::Circuit MimcCircuit() {
    circuit::Circuit c;
    c.resize(GKRB200_MIMC_LAYERS);
    c[2].In = {0};
    c[2].gate = gates::IdentityGate();
    for (int i = 0; i < 91; i++) {
        c[(size_t)i + 3].In = {2, i == 0 ? 1 : i + 2};
        c[(size_t)i + 3].gate = gates::NewCipherGate(Ark(i));
    }
    return circuit::BuildCircuit(c);
}
inline bool IsMimcCircuit(const circuit::Circuit& c) {
    if (c.size() != GKRB200_MIMC_LAYERS) return false;
    const circuit::Circuit m = MimcCircuit();
    for (size_t l = 0; l < c.size(); l++)
        if (c[l].In != m[l].In || c[l].gate.kind != m[l].gate.kind || (c[l].gate.kind == GKRB200_GATE_CIPHER && c[l].gate.ark != m[l].gate.ark)) return false;
    return true;
}
}  // namespace examples

namespace circuit {
// What the Go shim does before Assign / gkr.Prove: the description of the Circuit it was called with goes to
// gkrb200_check_mimc_circuit, which refuses (Panic, message names the layer) anything but examples.MimcCircuit().
inline void CheckIsMimc(const Circuit& c) {
    std::vector<int> n_in(c.size()), in, kinds(c.size());
    std::vector<fr::Element> arks(c.size());
    for (size_t l = 0; l < c.size(); l++) {
        n_in[l] = (int)c[l].In.size();
        in.insert(in.end(), c[l].In.begin(), c[l].In.end());
        kinds[l] = c[l].gate.kind;
        arks[l] = c[l].gate.ark;
    }
    in.push_back(0);
    check(gkrb200_check_mimc_circuit((int)c.size(), n_in.data(), in.data(), kinds.data(), c.empty() ? nullptr : arks[0].data()));
}
// Circuit.Assign(inputs...) for examples.MimcCircuit(): inputs[0] = key (layer 0), inputs[1] = message block (layer 1)
inline Assignment Assign(Device& d, const Circuit& c, const poly::MultiLin& key, const poly::MultiLin& msg) {
    CheckIsMimc(c);
    if (key.size() != msg.size()) throw Panic(GKRB200_ERR_ARG, "inconsistent input sizes");
    check(gkrb200_mimc_assign(d.handle(), key[0].data(), msg[0].data(), key.size(), nullptr));
    int bn = 0;
    while (((size_t)1 << bn) < key.size()) bn++;
    return Assignment(d, key.size(), bn);
}
}  // namespace circuit

namespace sumcheck {
using Proof = std::vector<std::vector<fr::Element>>;  // sumcheck/prover.go:22
// sumcheck.Prove(X, qPrimes, claims, gate) -> (proof, challenges, finalClaims)
inline std::tuple<Proof, std::vector<fr::Element>, std::vector<fr::Element>> Prove(Device& d, const std::vector<poly::MultiLin>& X,
                                                                                    const std::vector<std::vector<fr::Element>>& qPrimes,
                                                                                    const std::vector<fr::Element>& claims, const circuit::Gate& gate) {
    const int bn = (int)qPrimes.at(0).size();
    const size_t nco = (size_t)gate.Degree() + 2;
    std::vector<fr::Element> q, flat((size_t)bn * nco), challenges((size_t)bn), fin(1 + X.size());
    for (const auto& qp : qPrimes) {
        if ((int)qp.size() != bn) throw Panic(GKRB200_ERR_ARG, "inconsistent sizes of the qPrimes");
        q.insert(q.end(), qp.begin(), qp.end());
    }
    if (X.at(0).size() != ((size_t)1 << bn)) throw Panic(GKRB200_ERR_ARG, "inconsistent sizes : the table and qPrime disagree");  // prover.go:54
    check(gkrb200_sumcheck_prove(d.handle(), X[0][0].data(), X.size() > 1 ? X[1][0].data() : nullptr, bn, detail::ptr(q), qPrimes.size(),
                                 detail::ptr(claims), claims.size(), gate.kind, gate.kind == GKRB200_GATE_CIPHER ? gate.ark.data() : nullptr,
                                 detail::ptr(flat), detail::ptr(challenges), fin[0].data()));
    Proof proof((size_t)bn);
    for (int k = 0; k < bn; k++) proof[(size_t)k].assign(flat.begin() + (size_t)k * nco, flat.begin() + (size_t)(k + 1) * nco);
    return {proof, challenges, fin};
}
// sumcheck.Verify(claims, proof) -> (challenges, finalClaim, recombChal, err)   sumcheck/verifier.go:28-65
inline std::tuple<std::vector<fr::Element>, fr::Element, fr::Element, std::string> Verify(const std::vector<fr::Element>& claims, const Proof& proof) {
    const int bn = (int)proof.size();
    const size_t nco = bn ? proof[0].size() : 1;
    std::vector<fr::Element> flat, challenges((size_t)bn);
    for (const auto& round : proof) {
        if (round.size() != nco) throw Panic(GKRB200_ERR_ARG, "ragged sumcheck proof");
        flat.insert(flat.end(), round.begin(), round.end());
    }
    fr::Element fin{}, recomb{};
    const int rc = gkrb200_sumcheck_verify(detail::ptr(claims), claims.size(), detail::ptr(flat), bn, (int)nco,
                                           detail::ptr(challenges), fin.data(), recomb.data());
    if (rc == GKRB200_ERR_VERIFY) return {{}, {}, {}, gkrb200_last_error()};
    check(rc);
    return {challenges, fin, recomb, ""};
}
// sumcheck.Evaluation(gate, qPrime, claims, X...)   sumcheck/instance.go:49-68.  sum_x Eq(x) gate(X(x)), Eq = sum_j rho^j eq(q_j, .):
// the gate values are formed on the device and sum_x eq(q, x) g(x) is MultiLin.Evaluate(g, q), also on the device.
inline fr::Element Evaluation(Device& d, const circuit::Gate& gate, const std::vector<std::vector<fr::Element>>& qPrimes, const std::vector<fr::Element>& claims,
                              const std::vector<poly::MultiLin>& X) {
    const size_t n = X.at(0).size();
    poly::MultiLin g = X[0];
    if (gate.kind == GKRB200_GATE_CIPHER) {
        poly::MultiLin arks(n, gate.ark), t(n);
        check(gkrb200_fr_batch(d.handle(), 1, X[0][0].data(), X.at(1)[0].data(), n, t[0].data()));
        check(gkrb200_fr_batch(d.handle(), 1, t[0].data(), arks[0].data(), n, t[0].data()));
        check(gkrb200_fr_batch(d.handle(), 3, t[0].data(), nullptr, n, g[0].data()));
    }
    const size_t n_q = claims.empty() ? 1 : qPrimes.size();  // prover.go:117-119: without claims only qPrimes[0] is used
    fr::Element res = poly::Evaluate(d, g, qPrimes.at(0));
    if (n_q > 1) {
        const fr::Element rho = common::GetChallenge(claims);
        fr::Element pw = rho;
        for (size_t j = 1; j < n_q; j++) {
            res = fr::Add(res, fr::Mul(poly::Evaluate(d, g, qPrimes[j]), pw));
            pw = fr::Mul(pw, rho);
        }
    }
    return res;
}
}  // namespace sumcheck

namespace gkr {
struct Proof {  // gkr/prover.go:14-18
    std::vector<sumcheck::Proof> SumcheckProofs;
    std::vector<std::vector<fr::Element>> Claims;
    std::vector<std::vector<std::vector<fr::Element>>> QPrimes;
};
// gadget.GkrProofToVec (prover/gadget/hints.go:236-271): all SumcheckProofs[l][k][j], all Claims[l][j], all QPrimes[l][j][k]
inline std::vector<fr::Element> GkrProofToVec(const Proof& p) {
    std::vector<fr::Element> v;
    for (const auto& sp : p.SumcheckProofs)
        for (const auto& round : sp) v.insert(v.end(), round.begin(), round.end());
    for (const auto& cl : p.Claims) v.insert(v.end(), cl.begin(), cl.end());
    for (const auto& ql : p.QPrimes)
        for (const auto& q : ql) v.insert(v.end(), q.begin(), q.end());
    return v;
}
inline Proof ProofFromVec(const circuit::Circuit& c, int bn, const std::vector<fr::Element>& v) {
    Proof p;
    size_t cur = 0;
    const size_t L = c.size(), ubn = (size_t)bn;
    p.SumcheckProofs.resize(L);
    p.Claims.resize(L);
    p.QPrimes.resize(L);
    for (size_t l = 0; l < L; l++) {
        if (c[l].gate.nil()) continue;
        const size_t nco = (size_t)c[l].gate.Degree() + 2;
        p.SumcheckProofs[l].resize(ubn);
        for (size_t k = 0; k < ubn; k++, cur += nco) p.SumcheckProofs[l][k].assign(v.begin() + cur, v.begin() + cur + nco);
    }
    for (size_t l = 0; l < L; l++) {
        p.Claims[l].assign(v.begin() + cur, v.begin() + cur + c[l].Out.size());
        cur += c[l].Out.size();
    }
    for (size_t l = 0; l < L; l++) {
        const size_t nq = l == L - 1 ? 1 : c[l].Out.size();
        p.QPrimes[l].resize(nq);
        for (size_t j = 0; j < nq; j++, cur += ubn) p.QPrimes[l][j].assign(v.begin() + cur, v.begin() + cur + ubn);
    }
    if (cur != v.size()) throw Panic(GKRB200_ERR_STATE, "proof vector length mismatch");
    return p;
}
// gkr.Prove(c, a, qPrime)
inline Proof Prove(const circuit::Circuit& c, const circuit::Assignment& a, const std::vector<fr::Element>& qPrime) {
    circuit::CheckIsMimc(c);
    std::vector<fr::Element> v(gkrb200_proof_vec_len(a.bn()));
    check(gkrb200_gkr_prove_mimc(a.device().handle(), detail::ptr(qPrime), (int)qPrime.size(), v[0].data(), GKRB200_PROOF_MONTGOMERY));
    return ProofFromVec(c, a.bn(), v);
}
// gkr.Verify(c, proof, inputs, outputs, qPrime): inputs/outputs are the layers 0, 1 and 93 of the assignment held by the device.
// Returns "" when the proof is accepted, else the reason (Go: error).
inline std::string Verify(const circuit::Circuit&, const Proof& proof, const circuit::Assignment& a, const std::vector<fr::Element>& qPrime) {
    const std::vector<fr::Element> v = GkrProofToVec(proof);
    const int rc = gkrb200_gkr_verify_mimc(a.device().handle(), v[0].data(), a.bn(), detail::ptr(qPrime), GKRB200_PROOF_MONTGOMERY);
    if (rc == 0) return "";
    if (rc == GKRB200_ERR_VERIFY) return gkrb200_last_error();
    throw Panic(rc, gkrb200_last_error());
}
// gkr.Verify(c, proof, inputs, outputs, qPrime) with the CALLER's tables (gkr/verifier.go:15; the hint's self-check passes the solver's
// outputs, prover/gadget/hints.go:225-229): their MLEs are evaluated on the device from these bytes, not from the prover's assignment.
inline std::string Verify(Device& d, const circuit::Circuit& c, const Proof& proof, const std::vector<poly::MultiLin>& inputs, const poly::MultiLin& outputs,
                          const std::vector<fr::Element>& qPrime) {
    circuit::CheckIsMimc(c);
    if (inputs.size() != 2 || inputs[0].size() != outputs.size() || inputs[1].size() != outputs.size() || outputs.size() != ((size_t)1 << qPrime.size()))
        throw Panic(GKRB200_ERR_ARG, "inconsistent sizes of inputs / outputs / qPrime");
    const std::vector<fr::Element> v = GkrProofToVec(proof);
    const int rc = gkrb200_gkr_verify_mimc_io(d.handle(), v[0].data(), (int)qPrime.size(), detail::ptr(qPrime), GKRB200_PROOF_MONTGOMERY,
                                              inputs[0][0].data(), inputs[1][0].data(), outputs[0].data());
    if (rc == 0) return "";
    if (rc == GKRB200_ERR_VERIFY) return gkrb200_last_error();
    throw Panic(rc, gkrb200_last_error());
}
}  // namespace gkr

}  // namespace gkrmimc
