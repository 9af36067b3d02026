"""CPU tests of the oracle (oracle/pyref.py big-int restatement and oracle/gkr_oracle.c) against the reference's own
goldens and against each other.  No GPU.  The reference has exactly four pinned results on this path
(SURVEY.md section 8c); everything else it tests by round trip, which is mirrored here."""
import hashlib
import json
import os

import numpy as np
import pytest

import pyref as P

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
Q = P.Q


# ----------------------------------------------------------------- constants re-derived from q alone
def test_field_constants():
    assert Q == 21888242871839275222246405745257275088548364400416034343698204186575808495617
    assert Q.bit_length() == 254
    R = (1 << 256) % Q
    assert [hex((R >> (64 * i)) & (2**64 - 1)) for i in range(4)] == ["0xac96341c4ffffffb", "0x36fc76959f60cd29", "0x666ea36f7879462e", "0xe0a77c19a07df2f"]
    assert (-pow(Q, -1, 1 << 64)) % (1 << 64) == 0xC2E1F593EFFFFFFF
    assert (-pow(Q, -1, 1 << 32)) % (1 << 32) == 0xEFFFFFFF
    r2 = R * R % Q
    assert (r2 & (2**64 - 1)) == 0x1BB8E645AE216DA7


def test_fast_multiplier_equals_portable_and_bigint(oracle):
    """oracle/fr.h: the MULX/ADCX/ADOX product (the class of multiplier gnark-crypto's amd64 assembly is) against the textbook
    unsigned __int128 CIOS form and against Python big ints, edge values and random"""
    import random
    rnd = random.Random(5)
    R = (1 << 256) % Q
    rinv = pow(R, -1, Q)
    edge = [0, 1, 2, Q - 1, Q - 2, R, R * R % Q, Q // 2, Q // 2 + 1, (1 << 64) - 1, (1 << 128) - 1, (1 << 192) - 1, (1 << 253) - 1, Q - (1 << 64)]
    vals = edge + [rnd.randrange(Q) for _ in range(200)]

    def enc(v):
        return np.array([(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)], dtype=np.uint64)

    for x in edge + vals[-20:]:
        for y in vals:
            got = oracle.fr_mul(enc(x), enc(y))
            assert np.array_equal(got, oracle.fr_mul_portable(enc(x), enc(y)))
            assert np.array_equal(got, enc(x * y * rinv % Q))
    assert "CIOS" in oracle.fr_mul_kind()
    thr, lat = oracle.bench_fr_mul(20000)
    assert 0 < thr < 1e4 and 0 < lat < 1e4


# ----------------------------------------------------------------- the reference's goldens
def test_mimc_kat_reference_golden(oracle):
    """hash/hash_test.go:21-27 TestMimcCase"""
    want = 1808205620575546259657963589762746470347087906694759866517376279978241663265
    assert P.mimc_hash([12]) == want
    assert oracle.from_mont(oracle.mimc_hash(oracle.to_mont([12])))[0] == want
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    assert str(P.mimc_hash([0])) == kat["mimc_hash_0"]  # hashOfZeroes, prover/gadget/gadget.go:26-31
    assert oracle.from_mont(oracle.mimc_hash(oracle.to_mont([0])))[0] == int(kat["mimc_hash_0"])


def test_fold_reference_golden(oracle):
    """poly/multilin_test.go:12-31 TestFold: [0,1,2,3].Fold(5) == [10,11]"""
    assert P.fold([0, 1, 2, 3], 5) == [10, 11]
    assert oracle.from_mont(oracle.fold(oracle.to_mont([0, 1, 2, 3]), oracle.to_mont([5])[0])) == [10, 11]


def test_lagrange_reference_property(oracle):
    """poly/lagrange_test.go:10-29: the basis is 0/1 on the domain (pins Inverse)"""
    for d in range(1, 13):
        lag = oracle.lagrange_coefficient(d)
        plag = P.lagrange_coefficient(d)
        for i in range(d):
            assert oracle.from_mont(lag[i]) == plag[i]
            for x in range(d):
                assert P.eval_univariate(plag[i], x) == (1 if x == i else 0)


def test_univariate_reference_golden(oracle):
    """snark/polynomial/univariate_test.go:39-45: X^3+2X^2+3X+4 -> p(0)+p(1)=14, p(5)=194 (coefficients low->high)"""
    c = [4, 3, 2, 1]
    assert P.eval_univariate(c, 0) + P.eval_univariate(c, 1) == 14 and P.eval_univariate(c, 5) == 194
    cm = oracle.to_mont(c)
    assert oracle.from_mont(oracle.eval_univariate(cm, oracle.to_mont([5])[0]))[0] == 194


def test_random_fr_array(oracle):
    """common/common.go:49-55"""
    assert oracle.from_mont(oracle.random_fr_array(300)) == P.random_fr_array(300)
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    assert [str(x) for x in P.random_fr_array(4)] == kat["random_fr_array_4"]
    big = (1 << 33) + 5  # the uint64 product wraps
    assert P.random_fr_element(big) == ((big * big) % (1 << 64)) ^ 0xF45C9DF123F


# ----------------------------------------------------------------- C oracle vs big-int restatement
def _rand(rng, n):
    return [int.from_bytes(rng.bytes(40), "little") % Q for _ in range(n)]


def test_field_ops_c_vs_bigint(oracle):
    rng = np.random.default_rng(0)
    edge = [0, 1, 2, Q - 1, Q - 2, (1 << 256) % Q, Q // 2, (1 << 253), 2**64 - 1, 2**128 - 1]
    vals = edge + _rand(rng, 200)
    m = oracle.to_mont(vals)
    for i in range(0, len(vals)):
        j = (i * 7 + 3) % len(vals)
        a, b = vals[i], vals[j]
        assert oracle.from_mont(oracle.fr_mul(m[i], m[j]))[0] == a * b % Q
        assert oracle.from_mont(oracle.fr_add(m[i], m[j]))[0] == (a + b) % Q
        assert oracle.from_mont(oracle.fr_sub(m[i], m[j]))[0] == (a - b) % Q
    for v, mv in list(zip(vals, m))[:20]:
        assert oracle.from_mont(oracle.fr_inv(mv))[0] == (pow(v, Q - 2, Q))


@pytest.mark.parametrize("bn", range(0, 9))
def test_eq_tables(oracle, bn):
    """poly/eq_test.go:12-58: folded == chunked (every chunk size) and Evaluate(h) == EvalEq(q,h)"""
    rng = np.random.default_rng(bn)
    q = _rand(rng, bn)
    qm = oracle.to_mont(q).reshape(bn, 4)
    t = oracle.folded_eq_table(qm)
    assert oracle.from_mont(t) == P.folded_eq_table(q)
    for logc in range(0, bn + 1):
        assert np.array_equal(oracle.chunked_eq_table(qm, 1 << logc), t)
    h = _rand(rng, bn)
    hm = oracle.to_mont(h).reshape(bn, 4)
    assert oracle.from_mont(oracle.evaluate(t, hm))[0] == P.eval_eq(q, h) == P.evaluate(P.folded_eq_table(q), h)
    assert oracle.from_mont(oracle.eval_eq(qm, hm))[0] == P.eval_eq(q, h)
    mult = _rand(rng, 1)[0]
    assert oracle.from_mont(oracle.folded_eq_table(qm, oracle.to_mont([mult])[0])) == P.folded_eq_table(q, mult)


@pytest.mark.parametrize("kind", ["cipher", "identity"])
def test_partial_evals_and_interpolation(oracle, kind):
    rng = np.random.default_rng(5)
    bn = 5
    n = 1 << bn
    eq, L, Rr = _rand(rng, n), _rand(rng, n), _rand(rng, n)
    ark = _rand(rng, 1)[0]
    gate = P.Gate(kind, ark)
    X = [L, Rr] if kind == "cipher" else [L]
    ev = P.partial_evals(eq, X, gate)
    got = oracle.partial_evals(oracle.to_mont(eq), oracle.to_mont(L), oracle.to_mont(Rr) if kind == "cipher" else None,
                               oracle.GATE_CIPHER if kind == "cipher" else oracle.GATE_IDENTITY, oracle.to_mont([ark])[0])
    assert oracle.from_mont(got) == ev
    co = P.interpolate_on_range(ev)
    assert oracle.from_mont(oracle.interpolate_on_range(got)) == co
    for t in range(len(ev)):
        assert P.eval_univariate(co, t) == ev[t]


@pytest.mark.parametrize("bn", range(0, 7))
def test_sumcheck_cipher_c_vs_bigint_and_verifier(oracle, bn):
    """sumcheck/prover_test.go TestWithCipherGate on InitializeCipherGateInstance (testing.go:11-26)"""
    X, claims, qs, gate = P.init_cipher_gate_instance(bn)
    proof, chal, fin = P.sumcheck_prove(X, qs, claims, gate)
    cproof, cchal, cfin = oracle.sumcheck_prove([oracle.to_mont(X[0]), oracle.to_mont(X[1])], oracle.to_mont(qs[0]).reshape(1, bn, 4), oracle.to_mont(claims),
                                                oracle.GATE_CIPHER, oracle.to_mont([145646])[0])
    assert oracle.from_mont(cproof) == [x for r in proof for x in r]
    assert oracle.from_mont(cchal) == chal and oracle.from_mont(cfin) == fin
    vchal, vfinal, _ = P.sumcheck_verify(claims, proof)
    assert vchal == chal
    assert vfinal == gate.eval(fin[1], fin[2]) * fin[0] % Q  # prover_test.go:66-77
    rc, c2, f2, _ = oracle.sumcheck_verify(oracle.to_mont(claims), cproof)
    assert rc == 0 and oracle.from_mont(f2)[0] == vfinal


@pytest.mark.parametrize("bn,ninst", [(0, 10), (1, 10), (4, 10), (5, 91)])
def test_sumcheck_multi_identity_c_vs_bigint(oracle, bn, ninst):
    """sumcheck/prover_test.go TestWithMultiIdentity on InitializeMultiInstance (testing.go:28-57)"""
    X, claims, qs, gate = P.init_multi_instance(bn, ninst)
    proof, chal, fin = P.sumcheck_prove(X[:1], qs, claims, gate)
    qm = np.stack([oracle.to_mont(q).reshape(bn, 4) for q in qs])
    cproof, cchal, cfin = oracle.sumcheck_prove([oracle.to_mont(X[0])], qm, oracle.to_mont(claims), oracle.GATE_IDENTITY)
    assert oracle.from_mont(cproof) == [x for r in proof for x in r]
    assert oracle.from_mont(cfin) == fin
    # the random linear combination of the claims equals the brute-force evaluation (prover_test.go:48-57)
    eq, rho = P.make_eq_table(claims, qs)
    assert P.eval_univariate(claims, rho) == P.evaluation(gate, qs, claims, X[0])
    P.sumcheck_verify(claims, proof)


@pytest.mark.parametrize("bn", range(0, 5))
def test_gkr_c_vs_bigint_golden_and_verify(oracle, bn):
    """gkr/gkr_test.go:14-78 TestGKR (inputs :23-25) + committed digests"""
    gold = json.load(open(os.path.join(GOLDEN, "gkr_proof_digests.json")))
    c = P.mimc_circuit()
    blk, qp = P.random_fr_array(1 << bn), P.random_fr_array(bn)
    a = P.assign(c, blk, blk)
    pr = P.gkr_prove(c, a, qp)
    vec = P.gkr_proof_to_vec(pr)
    assert len(vec) == 1006 * bn + 183  # prover/gadget/hints.go:76-116
    assert hashlib.sha256(b"".join(x.to_bytes(32, "big") for x in vec)).hexdigest() == gold[str(bn)]
    assert P.gkr_verify(c, pr, [blk, blk], a[93], qp)
    bm, qm = oracle.to_mont(blk), oracle.to_mont(qp).reshape(bn, 4)
    out93, cvec = oracle.assign_and_prove_mimc(bm, bm, qm)
    assert oracle.from_mont(cvec) == vec
    assert oracle.from_mont(out93) == a[93]
    assert oracle.gkr_verify_mimc(cvec, bm, bm, out93, qm) == 0
    # claims agree with MLE evaluations of an untouched assignment (gkr_test.go:35-45)
    for layer in (0, 1, 2, 50):
        for j in range(0, len(pr.Claims[layer]), 30):
            assert P.evaluate(a[layer], pr.QPrimes[layer][j]) == pr.Claims[layer][j]
    bad = cvec.copy()
    bad[len(bad) // 3, 1] ^= np.uint64(4)
    assert oracle.gkr_verify_mimc(bad, bm, bm, out93, qm) != 0


def test_circuit_structure(oracle):
    """examples/mimc_test.go:19-54: arity, sorted Out lists; keyed permutation on element 0"""
    c = P.mimc_circuit()
    assert len(c) == 94 and not c[0].In and not c[1].In
    for lay in c:
        assert lay.Out == sorted(lay.Out)
    assert c[2].Out == list(range(3, 94)) and c[0].Out == [2] and c[1].Out == [3] and c[93].Out == []
    key, msg = P.random_fr_array(8), P.random_fr_array(8)[::-1]
    a = P.assign(c, key, msg)
    assert a[93][0] == P.mimc_keyed_permutation(msg[0], key[0])
    lay = oracle.mimc_assign(oracle.to_mont(key), oracle.to_mont(msg))
    for l in (2, 3, 47, 93):
        assert oracle.from_mont(lay[l]) == a[l]


def test_c_oracle_worker_pool_has_no_stale_ticket_race(oracle):
    """Regression: the worker pool once validated a ticket drawn in an earlier run against the NEXT run's task count, so under
    preemption a task ran twice and the caller went on one task early -- ~3 % of the 91-claim eq tables (181 back-to-back tiny
    pool runs, sumcheck/prover.go:121-141) came out wrong at bn = 11.  200 oversubscribed repetitions catch that with
    probability > 0.99; the oracle is the checker of every GPU parity test, so it must be exactly reproducible."""
    bn, nq = 11, 91
    qp = oracle.random_fr_array(nq * bn + 7)[7:].reshape(nq, bn, 4)
    cl = oracle.random_fr_array(nq + 3)[3:]
    oracle.set_threads(1)
    ref_eq, ref_rho = oracle.make_eq_table(cl, qp)
    try:
        for threads in (8, 32):
            oracle.set_threads(threads)
            for _ in range(100):
                eq, rho = oracle.make_eq_table(cl, qp)
                assert np.array_equal(rho, ref_rho)
                assert np.array_equal(eq, ref_eq), "the multi-threaded eq table differs from the single-threaded one"
    finally:
        oracle.set_threads(min(8, os.cpu_count() or 1))


def test_c_oracle_thread_count_invariance(oracle):
    """the worker-pool decomposition (common/parallelize.go, sumcheck/worker.go) must not change any word"""
    bn = 11
    key, qp = oracle.random_fr_array(1 << bn), oracle.random_fr_array(bn)
    res = []
    for t in (1, 3, 8):
        oracle.set_threads(t)
        res.append(oracle.assign_and_prove_mimc(key, key, qp)[1])
    oracle.set_threads(min(8, os.cpu_count() or 1))
    assert np.array_equal(res[0], res[1]) and np.array_equal(res[0], res[2])
