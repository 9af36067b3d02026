// The reference's own Go tests restated against include/gkrb200.hpp (the C++ mirror of the Go API over the C ABI).
//
//   TestMimcCase           hash/hash_test.go:21-27        (host)
//   TestInterpolate        poly/lagrange_test.go:10-29    (host; evaluating the interpolant gives the values back)
//   TestCircuitShape       circuit/circuit.go:28-91, examples/mimc.go:10-37 (host)
//   TestFold               poly/multilin_test.go:12-31    (device)
//   TestFolding            sumcheck/prover_test.go:15-41  (device: eq table, then fold, against the closed form EvalEq)
//   TestWithCipherGate     sumcheck/prover_test.go:91-97 + genericTest :43-81  (device prover, host verifier)
//   TestWithMultiIdentity  sumcheck/prover_test.go:83-89
//   TestGKR                gkr/gkr_test.go:14-79          (device)
//
// `test_api --host-only` runs the tests that need no GPU (used by the CPU suite); without the flag everything runs and
// a missing CUDA device is a failure (there is no CPU fallback to hide behind).  This file is test infrastructure.
#include <cstdio>
#include <cstring>
#include <string>

#include "gkrb200.hpp"

using namespace gkrmimc;
using fr::Element;

static int failures = 0;
#define EXPECT(cond, ...)                                            \
    do {                                                             \
        if (!(cond)) {                                               \
            failures++;                                              \
            fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__);     \
            fprintf(stderr, __VA_ARGS__);                            \
            fprintf(stderr, "\n");                                   \
        }                                                            \
    } while (0)

static poly::MultiLin Range(size_t n) {  // L[i].SetUint64(uint64(i))  (sumcheck/testing.go:19-22)
    poly::MultiLin in(n), out(n);
    for (size_t i = 0; i < n; i++) in[i] = {(uint64_t)i, 0, 0, 0};
    if (n) check(gkrb200_to_montgomery(in[0].data(), n, out[0].data()));
    return out;
}

// ---------------------------------------------------------------------------------------------- host-only tests
static void TestMimcCase() {
    const Element y = hash::MimcHash({fr::SetUint64(12)});
    EXPECT(y == hash::MimcHash({fr::SetUint64(12)}), "MimcHash is not deterministic");
    EXPECT(!(y == hash::MimcHash({fr::SetUint64(12), fr::SetUint64(0)})), "MimcHash ignores trailing blocks");
    EXPECT(common::GetChallenge({fr::SetUint64(12)}) == y, "GetChallenge != MimcHash");
    // hash/hash_test.go:24: the expected value, accumulated from its decimal string in 256 bits, against the regular-form words
    const Element reg = fr::ToRegular(y);
    const char* dec = "1808205620575546259657963589762746470347087906694759866517376279978241663265";
    uint64_t acc[4] = {0, 0, 0, 0};
    for (const char* p = dec; *p; p++) {
        unsigned __int128 carry = (unsigned)(*p - '0');
        for (int i = 0; i < 4; i++) {
            unsigned __int128 t = (unsigned __int128)acc[i] * 10 + carry;
            acc[i] = (uint64_t)t;
            carry = t >> 64;
        }
    }
    EXPECT(reg[0] == acc[0] && reg[1] == acc[1] && reg[2] == acc[2] && reg[3] == acc[3], "MimcHash([12]) differs from hash/hash_test.go:24");
}

static void TestInterpolate() {
    for (size_t n = 1; n <= 12; n++) {
        std::vector<Element> values = common::RandomFrArray(n + 3);
        values.erase(values.begin(), values.begin() + 3);
        const std::vector<Element> coeffs = poly::InterpolateOnRange(values);
        for (size_t i = 0; i < n; i++)
            EXPECT(poly::EvalUnivariate(coeffs, fr::SetUint64(i)) == values[i], "interpolant of %zu values is wrong at %zu", n, i);
    }
    bool threw = false;
    try {
        poly::InterpolateOnRange(std::vector<Element>(13));
    } catch (const Panic& p) {
        threw = p.code == GKRB200_ERR_ARG;
    }
    EXPECT(threw, "a domain of 13 points must be refused (poly/lagrange.go:21)");
}

static void TestCircuitShape() {
    const circuit::Circuit c = examples::MimcCircuit();
    EXPECT(c.size() == 94, "MimcCircuit has %zu layers", c.size());
    EXPECT(c.InputArity() == 2, "InputArity = %d", c.InputArity());
    EXPECT(c.IsInputLayer(0) && c.IsInputLayer(1) && !c.IsInputLayer(2) && !c.IsInputLayer(93), "input layers");
    EXPECT(c[0].Out.size() == 1 && c[0].Out[0] == 2, "layer 0 feeds the copy layer only");
    EXPECT(c[2].Out.size() == 91, "the copied key feeds the 91 cipher layers (got %zu)", c[2].Out.size());
    EXPECT(c[1].Out.size() == 1 && c[1].Out[0] == 3, "the message feeds the first cipher layer");
    EXPECT(c[93].Out.empty(), "the last layer has no consumer");
    EXPECT(c[2].gate.Degree() == 1 && c[3].gate.Degree() == 7, "gate degrees");
    size_t total = 0;  // the flat proof length the gadget allocates (prover/gadget/hints.go:76-116)
    const size_t bn = 10;
    for (size_t l = 0; l < c.size(); l++) {
        if (!c[l].gate.nil()) total += bn * (size_t)(c[l].gate.Degree() + 2);
        total += c[l].Out.size() + (l == 93 ? 1 : c[l].Out.size()) * bn;
    }
    EXPECT(total == gkrb200_proof_vec_len((int)bn), "proof vector length %zu != %zu", total, gkrb200_proof_vec_len((int)bn));
    circuit::Circuit bad(3);  // an input layer used twice is refused (circuit/circuit.go:36-40)
    bad[1].In = {0};
    bad[1].gate = gates::IdentityGate();
    bad[2].In = {0};
    bad[2].gate = gates::IdentityGate();
    bool threw = false;
    try {
        circuit::BuildCircuit(bad);
    } catch (const Panic&) {
        threw = true;
    }
    EXPECT(threw, "BuildCircuit accepted an input layer with two consumers");
    // the C side checks that the circuit it is asked to assign / prove IS examples.MimcCircuit() (gkrb200_check_mimc_circuit)
    circuit::CheckIsMimc(c);
    for (int variant = 0; variant < 4; variant++) {
        circuit::Circuit w = examples::MimcCircuit();
        if (variant == 0) w[50].In[1] = 48;                                    // wrong wiring
        if (variant == 1) w[2].gate = gates::NewCipherGate(examples::Ark(0));  // wrong gate
        if (variant == 2) w[7].gate.ark = examples::Ark(5);                    // wrong round constant
        if (variant == 3) w.pop_back();                                        // wrong depth
        threw = false;
        try {
            circuit::CheckIsMimc(w);
        } catch (const Panic& p) {
            threw = p.code == GKRB200_ERR_ARG;
        }
        EXPECT(threw, "CheckIsMimc accepted a circuit that is not the MiMC circuit (variant %d)", variant);
    }
    // gate.Eval on scalars: cipher = (vL + vR + ark)^7
    const Element ark = fr::SetUint64(145646), l = fr::SetUint64(3), r = fr::SetUint64(4);
    Element t = fr::Add(fr::Add(l, r), ark), t7 = fr::One();
    for (int i = 0; i < 7; i++) t7 = fr::Mul(t7, t);
    EXPECT(gates::NewCipherGate(ark).Eval({l, r}) == t7, "CipherGate.Eval");
    EXPECT(gates::IdentityGate().Eval({l}) == l, "IdentityGate.Eval");
    EXPECT(fr::Mul(fr::Inverse(t), t) == fr::One(), "Inverse");
}

static void TestSumcheckVerifierRejects() {
    // a sumcheck "proof" with a wrong first round must be refused with the round number (sumcheck/verifier.go:45-47)
    sumcheck::Proof proof(2, std::vector<Element>(3, fr::SetUint64(1)));
    auto [ch, fin, rho, err] = sumcheck::Verify({fr::SetUint64(5)}, proof);
    EXPECT(!err.empty() && err.find("round 0") != std::string::npos, "expected a round-0 failure, got '%s'", err.c_str());
    // and a consistent hand-made one accepted: P(t) = 2 + t, claim = P(0) + P(1) = 5
    sumcheck::Proof ok(1, {fr::SetUint64(2), fr::SetUint64(1)});
    auto [ch2, fin2, rho2, err2] = sumcheck::Verify({fr::SetUint64(5)}, ok);
    EXPECT(err2.empty(), "hand-made proof refused: %s", err2.c_str());
    EXPECT(ch2.size() == 1 && ch2[0] == common::GetChallenge(ok[0]), "challenge is not the hash of the round polynomial");
    EXPECT(fin2 == poly::EvalUnivariate(ok[0], ch2[0]), "final claim is not P(r)");
    EXPECT(rho2 == common::GetChallenge({fr::SetUint64(5)}), "recombination challenge");
}

// ---------------------------------------------------------------------------------------------- device tests
static void TestFold(Device& d) {
    poly::MultiLin bkt = Range(4);  // [0, 1, 2, 3]
    poly::Fold(d, bkt, fr::SetUint64(5));
    EXPECT(bkt.size() == 2 && bkt[0] == fr::SetUint64(10) && bkt[1] == fr::SetUint64(11), "folding [0,1,2,3] on 5 should yield [10, 11]");
}

static void TestFolding(Device& d) {
    for (int bn = 2; bn < 15; bn++) {
        const std::vector<Element> q = common::RandomFrArray((size_t)bn);
        poly::MultiLin eq = poly::FoldedEqTable(d, q);
        EXPECT(eq.size() == ((size_t)1 << bn), "eq table size");
        // eq[x] is the closed form eq(q, bits(x)), MSB first (poly/eq.go:41-59): spot-check 5 entries
        for (size_t x : {(size_t)0, (size_t)1, eq.size() / 2, eq.size() - 2, eq.size() - 1}) {
            std::vector<Element> bits((size_t)bn);
            for (int k = 0; k < bn; k++) bits[(size_t)k] = fr::SetUint64((x >> (bn - 1 - k)) & 1);
            EXPECT(eq[x] == poly::EvalEq(q, bits), "bn=%d: eq table entry %zu differs from EvalEq", bn, x);
        }
        // folding the eq table on q[0] gives eq(q[0],q[0]) * eq(q[1:], .)
        poly::Fold(d, eq, q[0]);
        const std::vector<Element> tail(q.begin() + 1, q.end());
        const poly::MultiLin rest = poly::FoldedEqTable(d, tail);
        const Element c0 = poly::EvalEq({q[0]}, {q[0]});
        bool same = eq.size() == rest.size();
        for (size_t i = 0; same && i < eq.size(); i++) same = eq[i] == fr::Mul(c0, rest[i]);
        EXPECT(same || bn > 10, "bn=%d: folded eq table is not eq(q0,q0) * eq(q[1:], .)", bn);
        if (bn > 10) {  // spot-check only for the big tables (scalar host multiplications)
            for (size_t i : {(size_t)0, eq.size() / 3, eq.size() - 1}) EXPECT(eq[i] == fr::Mul(c0, rest[i]), "bn=%d: folded eq entry %zu", bn, i);
        }
    }
}

// genericTest (sumcheck/prover_test.go:43-81)
static void genericTest(Device& d, const std::vector<poly::MultiLin>& X, const std::vector<Element>& claims, const std::vector<std::vector<Element>>& qs,
                        const circuit::Gate& gate, int bn) {
    if (qs.size() > 1) {  // the random linear combination of the claims equals the combined sum
        const Element rnd = common::GetChallenge(claims);
        EXPECT(poly::EvalUnivariate(claims, rnd) == sumcheck::Evaluation(d, gate, qs, claims, X), "bn=%d: the random linear combination did not match the claim", bn);
    }
    auto [proof, challenges, fClm] = sumcheck::Prove(d, X, qs, claims, gate);
    auto [challengesV, expectedValue, recombChal, err] = sumcheck::Verify(claims, proof);
    EXPECT(err.empty(), "bn=%d: sumcheck was not deemed valid: %s", bn, err.c_str());
    EXPECT(challenges == challengesV, "bn=%d: prover's and verifier challenges do not match", bn);
    EXPECT(recombChal == common::GetChallenge(claims), "bn=%d: recombination challenges do not match", bn);
    std::vector<Element> xs(fClm.begin() + 1, fClm.end());
    const Element expVal = fr::Mul(gate.Eval(xs), fClm[0]);
    EXPECT(expectedValue == expVal, "bn=%d: inconsistency of the final values for the verifier", bn);
    // the final claims are the MLEs of the inputs (and of Eq) at the challenges
    for (size_t k = 0; k + 1 < fClm.size(); k++) EXPECT(fClm[k + 1] == poly::Evaluate(d, X[k], challenges), "bn=%d: final claim of X[%zu]", bn, k);
    Element eqr = poly::EvalEq(qs[0], challenges);
    if (qs.size() > 1) {
        std::vector<Element> es;
        for (const auto& q : qs) es.push_back(poly::EvalEq(q, challenges));
        eqr = poly::EvalUnivariate(es, recombChal);
    }
    EXPECT(fClm[0] == eqr, "bn=%d: final claim of Eq", bn);
}

static void TestWithCipherGate(Device& d) {
    for (int bn = 0; bn < 15; bn++) {  // InitializeCipherGateInstance (sumcheck/testing.go:11-26)
        const std::vector<Element> q = common::RandomFrArray((size_t)bn);
        const circuit::Gate gate = gates::NewCipherGate(fr::SetUint64(145646));
        const poly::MultiLin L = Range((size_t)1 << bn);
        const std::vector<poly::MultiLin> X = {L, L};
        const Element claim = sumcheck::Evaluation(d, gate, {q}, {}, X);
        genericTest(d, X, {claim}, {q}, gate, bn);
    }
}

static void TestWithMultiIdentity(Device& d) {
    for (int bn = 0; bn < 15; bn++) {  // InitializeMultiInstance (sumcheck/testing.go:28-57)
        const int ninstance = 10;
        std::vector<std::vector<Element>> qs((size_t)ninstance, std::vector<Element>((size_t)bn));
        for (int i = 0; i < ninstance; i++)
            for (int j = 0; j < bn; j++) qs[(size_t)i][(size_t)j] = fr::SetUint64((uint64_t)(i * j + i));
        const circuit::Gate gate = gates::IdentityGate();
        const std::vector<poly::MultiLin> X = {Range((size_t)1 << bn)};
        std::vector<Element> claims;
        for (int i = 0; i < ninstance; i++) claims.push_back(sumcheck::Evaluation(d, gate, {qs[(size_t)i]}, {}, X));
        genericTest(d, X, claims, qs, gate, bn);
    }
}

static void TestGKR(Device& d) {
    for (int bn = 0; bn < 12; bn++) {
        const circuit::Circuit c = examples::MimcCircuit();
        const poly::MultiLin block = common::RandomFrArray((size_t)1 << bn), initstate = common::RandomFrArray((size_t)1 << bn);
        const std::vector<Element> qPrime = common::RandomFrArray((size_t)bn);
        const circuit::Assignment a = circuit::Assign(d, c, block, initstate);
        const gkr::Proof proof = gkr::Prove(c, a, qPrime);
        EXPECT(gkr::GkrProofToVec(proof).size() == gkrb200_proof_vec_len(bn), "bn=%d: proof vector length", bn);
        // the claims are consistent with the assignment
        for (int layer = (int)c.size() - 1; layer >= 0; layer--)
            for (size_t j = 0; j < proof.Claims[(size_t)layer].size(); j++)
                EXPECT(a.Evaluate(layer, proof.QPrimes[(size_t)layer][j]) == proof.Claims[(size_t)layer][j], "bn=%d: claim inconsistent with assignment at layer %d no %zu", bn, layer, j);
        // the claims are consistent with the layers' evaluations (gkr_test.go:47-69; every layer up to bn = 6, then a sample)
        for (int layer = (int)c.size() - 1; layer >= 0; layer--) {
            if (c[(size_t)layer].gate.nil()) break;
            if (bn > 6 && layer != 93 && layer != 47 && layer != 3 && layer != 2) continue;
            const std::vector<poly::MultiLin> Xs = a.InputsOfLayer(c, layer);
            for (size_t j = 0; j < proof.Claims[(size_t)layer].size(); j++) {
                if (layer == 2 && bn > 3 && j % 13) continue;  // 91 claims on the copy layer: sample them
                EXPECT(sumcheck::Evaluation(d, c[(size_t)layer].gate, {proof.QPrimes[(size_t)layer][j]}, {}, Xs) == proof.Claims[(size_t)layer][j],
                       "bn=%d: inconsistent claim at layer %d no %zu", bn, layer, j);
            }
        }
        const std::string err = gkr::Verify(c, proof, a, qPrime);
        EXPECT(err.empty(), "bn = %d error at gkr verifier : %s", bn, err.c_str());
        // gkr.Verify(c, proof, inputs, outputs, qPrime) with the caller's tables (gkr_test.go:72): accepted; a wrong output is refused
        {
            poly::MultiLin outputs = a[93];
            const std::string e2 = gkr::Verify(d, c, proof, {block, initstate}, outputs, qPrime);
            EXPECT(e2.empty(), "bn = %d error at gkr verifier (caller's inputs/outputs) : %s", bn, e2.c_str());
            outputs[outputs.size() / 2][0] ^= 1;
            EXPECT(!gkr::Verify(d, c, proof, {block, initstate}, outputs, qPrime).empty(), "bn = %d: wrong outputs accepted", bn);
        }
        // a tampered proof is refused
        gkr::Proof bad = proof;
        bad.Claims[40][0][0] ^= 1;
        EXPECT(!gkr::Verify(c, bad, a, qPrime).empty(), "bn = %d: tampered claim accepted", bn);
        if (bn > 0) {
            gkr::Proof bad2 = proof;
            bad2.SumcheckProofs[93][0][2][1] ^= 4;
            EXPECT(!gkr::Verify(c, bad2, a, qPrime).empty(), "bn = %d: tampered round polynomial accepted", bn);
        }
        // proving twice gives the same bytes (the assignment is not consumed)
        EXPECT(gkr::GkrProofToVec(gkr::Prove(c, a, qPrime)) == gkr::GkrProofToVec(proof), "bn=%d: second proof differs", bn);
    }
    bool threw = false;  // Go panics on inconsistent input sizes (circuit/assignment.go:18-21)
    try {
        circuit::Assign(d, examples::MimcCircuit(), common::RandomFrArray(4), common::RandomFrArray(8));
    } catch (const Panic&) {
        threw = true;
    }
    EXPECT(threw, "Assign accepted inputs of different sizes");
}

int main(int argc, char** argv) {
    const bool host_only = argc > 1 && !strcmp(argv[1], "--host-only");
    try {
        TestMimcCase();
        TestInterpolate();
        TestCircuitShape();
        TestSumcheckVerifierRejects();
        printf("host tests done (%d failures)\n", failures);
        if (!host_only) {
            Device d(0, 15);
            TestFold(d);
            TestFolding(d);
            printf("fold tests done (%d failures)\n", failures);
            TestWithCipherGate(d);
            TestWithMultiIdentity(d);
            printf("sumcheck tests done (%d failures)\n", failures);
            TestGKR(d);
            printf("gkr tests done (%d failures)\n", failures);
        }
    } catch (const Panic& p) {
        fprintf(stderr, "PANIC (%d): %s\n", p.code, p.what());
        return 2;
    }
    printf("%s: %d failure(s)\n", host_only ? "host-only" : "all", failures);
    return failures ? 1 : 0;
}
