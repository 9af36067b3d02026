"""Pure-Python big-int restatement of the gkr-mimc prover/verifier hot path.

TEST INFRASTRUCTURE ONLY (oracle): imported by tests/, __graft_entry__.smoke()
and the golden-vector generator. The product (gkr-mimc_b200/) never imports it.

Values are *regular-form* integers mod q (no Montgomery): every reference
operation is exact field arithmetic on canonical elements, so the regular-form
result equals FromMont() of what the Go code holds.  Each function cites the
reference file:line it follows (paths relative to the reference root).

The third-party field gnark-crypto v0.6.1-0.20220110145513-493bb1c180d9
(ecc/bn254/fr, go.mod:7) is not vendored in the reference; its semantics are
restated here as "integers mod q" (SURVEY.md Appendix A).

Pinned against the reference's own goldens in tests/test_oracle.py:
  hash/hash_test.go:21-27, poly/multilin_test.go:12-31,
  poly/lagrange_test.go:10-29, snark/polynomial/univariate_test.go:39-45.
"""
import json
import os

Q = 21888242871839275222246405745257275088548364400416034343698204186575808495617
R = (1 << 256) % Q
MIMC_ROUNDS = 91  # hash/mimc.go:8

_here = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(_here, "..", "tests", "golden", "arks.json")) as _f:
    ARKS = [int(s) for s in json.load(_f)]  # hash/ark.go:232-335, first 91


# ---------------------------------------------------------------- hash/mimc.go
def sbox(x):
    """hash/poseidon.go:129-135  x^7 as ((x^2*x)^2)*x"""
    t = x
    x = x * x % Q
    x = x * t % Q
    x = x * x % Q
    return x * t % Q


def mimc_keyed_permutation(x, key):
    """hash/mimc.go:31-39"""
    res = x
    for i in range(MIMC_ROUNDS):
        res = sbox((res + key + ARKS[i]) % Q)
    return res


def mimc_block_cipher(msg, key):
    """hash/mimc.go:43-49"""
    return (mimc_keyed_permutation(msg, key) + key) % Q


def mimc_hash(inputs):
    """hash/mimc.go:11-28 (Miyaguchi-Preneel)"""
    state = 0
    for b in inputs:
        state = (state + mimc_block_cipher(b, state) + b) % Q
    return state


def get_challenge(seed):
    """common/challenge.go:10-12"""
    return mimc_hash(seed)


def random_fr_element(i):
    """common/common.go:52: SetUint64(uint64(i)*uint64(i) ^ 0xf45c9df123f) -- the product wraps at 2^64"""
    return (((i * i) & 0xFFFFFFFFFFFFFFFF) ^ 0xF45C9DF123F) % Q


def random_fr_array(n):
    """common/common.go:49-55"""
    return [random_fr_element(i) for i in range(n)]


# ------------------------------------------------------------------ poly/
def fold(tab, r):
    """poly/multilin.go:19-36  MSB-first: pairs (i, i+mid)"""
    mid = len(tab) // 2
    return [(tab[i] + r * (tab[i + mid] - tab[i])) % Q for i in range(mid)]


def evaluate(tab, coords):
    """poly/multilin.go:59-66"""
    t = list(tab)
    for r in coords:
        t = fold(t, r)
    return t[0]


def eval_eq(q, h):
    """poly/eq.go:19-32"""
    res = 1
    for a, b in zip(q, h):
        res = res * ((1 + 2 * a * b - a - b) % Q) % Q
    return res


def folded_eq_table(q, multiplier=1):
    """poly/eq.go:41-59"""
    n = len(q)
    t = [0] * (1 << n)
    t[0] = multiplier % Q
    for i, r in enumerate(q):
        for j in range(1 << i):
            J = j << (n - i)
            JN = J + (1 << (n - 1 - i))
            t[JN] = r * t[J] % Q
            t[J] = (t[J] - t[JN]) % Q
    return t


def chunk_of_eq_table(out, chunk_id, chunk_size, q, multiplier=1):
    """poly/eq.go:62-89"""
    n_chunks = (1 << len(q)) // chunk_size
    log_n = (n_chunks - 1).bit_length()  # common.Log2Ceil
    r = multiplier % Q
    for k in range(log_n):
        rho = q[log_n - k - 1]
        if (chunk_id >> k) & 1:
            r = r * rho % Q
        else:
            r = r * (1 - rho) % Q
    out[chunk_id * chunk_size:(chunk_id + 1) * chunk_size] = folded_eq_table(q[log_n:], r)


def eval_univariate(coeffs, x):
    """poly/lagrange.go:31-39 (coefficients low -> high)"""
    res = coeffs[-1]
    for c in reversed(coeffs[:-1]):
        res = (res * x + c) % Q
    return res


def lagrange_coefficient(domain):
    """poly/lagrange.go:42-92"""
    result = []
    for l in range(domain):
        acc = [0] * domain
        if domain:
            acc[0] = 1
        for i in range(domain):
            if i == l:
                continue
            upd = [0] * domain
            for j in range(domain):
                for k in range(min(2, domain - j)):
                    b = (-i) % Q if k == 0 else 1
                    upd[j + k] = (upd[j + k] + acc[j] * b) % Q
            acc = upd
        norm = eval_univariate(acc, l) if domain else 0
        inv = pow(norm, Q - 2, Q)  # fr.Inverse (0 -> 0)
        result.append([c * inv % Q for c in acc])
    return result


_LAGRANGE = {}


def interpolate_on_range(values):
    """poly/lagrange.go:96-111"""
    n = len(values)
    if n not in _LAGRANGE:
        _LAGRANGE[n] = lagrange_coefficient(n)
    lag = _LAGRANGE[n]
    res = [0] * n
    for i, v in enumerate(values):
        for j, c in enumerate(lag[i]):
            res[j] = (res[j] + c * v) % Q
    return res


# --------------------------------------------------------------- circuit/
class Gate:
    """circuit/gates.go:9-21; kind 'identity' (gates/copy.go) or 'cipher' (gates/cipher.go)"""

    def __init__(self, kind, ark=0):
        self.kind, self.ark = kind, ark % Q

    def degree(self):
        return 7 if self.kind == "cipher" else 1  # cipher.go:68-70, copy.go:30-32

    def eval(self, *xs):
        if self.kind == "cipher":  # cipher.go:45-55
            return sbox((xs[1] + self.ark + xs[0]) % Q)
        return xs[0]  # copy.go:20-22


class Layer:
    def __init__(self, ins, gate=None):
        self.In, self.Out, self.Gate = list(ins), [], gate


def mimc_circuit():
    """examples/mimc.go:10-37 + circuit/circuit.go:28-44"""
    c = [Layer([]), Layer([]), Layer([0], Gate("identity"))]
    for i in range(91):
        inp = 1 if i == 0 else i + 2
        c.append(Layer([2, inp], Gate("cipher", ARKS[i])))
    for l, lay in enumerate(c):
        for pos in lay.In:
            c[pos].Out.append(l)
    for l, lay in enumerate(c):
        assert not (len(lay.In) == 0 and len(lay.Out) > 1)
    return c


def assign(c, *inps):
    """circuit/assignment.go:12-32"""
    a = [list(x) for x in inps]
    for i in range(len(inps), len(c)):
        ins = [a[p] for p in c[i].In]
        a.append([c[i].Gate.eval(*[t[x] for t in ins]) for x in range(len(ins[0]))])
    return a


# --------------------------------------------------------------- sumcheck/
def make_eq_table(claims, qprimes):
    """sumcheck/prover.go:102-144; returns (eq, rho)"""
    assert not (len(claims) != len(qprimes) and len(qprimes) > 1)
    eq = folded_eq_table(qprimes[0])
    if len(claims) < 1:
        return eq, 0
    rho = get_challenge(claims)
    mult = rho
    for i in range(1, len(qprimes)):
        tmp = folded_eq_table(qprimes[i], mult)
        eq = [(a + b) % Q for a, b in zip(eq, tmp)]
        mult = mult * rho % Q
    return eq, rho


def partial_evals(eq, X, gate):
    """sumcheck/algo.go:54-205: evals[t] = sum_x eq_t(x) * gate(X_t(x)) for t = 0..degree"""
    n_evals = gate.degree() + 2
    mid = len(eq) // 2
    evals = [0] * n_evals
    for x in range(mid):
        e0, de = eq[x], (eq[x + mid] - eq[x]) % Q
        v0 = [t[x] for t in X]
        dv = [(t[x + mid] - t[x]) % Q for t in X]
        for t in range(n_evals):
            e = (e0 + t * de) % Q
            vs = [(a + t * d) % Q for a, d in zip(v0, dv)]
            evals[t] = (evals[t] + e * gate.eval(*vs)) % Q
    return evals


def sumcheck_prove(X, qprimes, claims, gate):
    """sumcheck/prover.go:46-90 -> (proof, challenges, final_claims)"""
    bn = len(qprimes[0])
    for x in X:
        assert len(x) == 1 << bn
    X = [list(x) for x in X]
    eq, _ = make_eq_table(claims, qprimes)
    proof, challenges = [], []
    for _ in range(bn):
        evals = partial_evals(eq, X, gate)
        coeffs = interpolate_on_range(evals)
        r = get_challenge(coeffs)
        eq = fold(eq, r)
        X = [fold(x, r) for x in X]
        proof.append(coeffs)
        challenges.append(r)
    return proof, challenges, [eq[0]] + [x[0] for x in X]


def sumcheck_verify(claims, proof):
    """sumcheck/verifier.go:28-65 -> (challenges, final_claim, recomb_chal); raises on failure"""
    rho = get_challenge(claims)
    expected = eval_univariate(claims, rho)
    challenges = []
    for i, p in enumerate(proof):
        actual = (eval_univariate(p, 0) + eval_univariate(p, 1)) % Q
        if actual != expected:
            raise ValueError("round %d: p(0)+p(1) mismatch" % i)
        r = get_challenge(p)
        challenges.append(r)
        expected = eval_univariate(p, r)
    return challenges, expected, rho


def evaluation(gate, qprimes, claims, *X):
    """sumcheck/instance.go:49-68 (brute force, tests only)"""
    eq, _ = make_eq_table(claims, qprimes)
    res = 0
    for n in range(len(X[0])):
        res = (res + gate.eval(*[x[n] for x in X]) * eq[n]) % Q
    return res


def init_cipher_gate_instance(bn):
    """sumcheck/testing.go:11-26"""
    q = random_fr_array(bn)
    gate = Gate("cipher", 145646)
    L = list(range(1 << bn))
    Rr = list(range(1 << bn))
    claim = evaluation(gate, [q], [], L, Rr)
    return [L, Rr], [claim], [q], gate


def init_multi_instance(bn, ninstance):
    """sumcheck/testing.go:28-57"""
    gate = Gate("identity")
    qs = [[(i * j + i) % Q for j in range(bn)] for i in range(ninstance)]
    L = list(range(1 << bn))
    Rr = list(range(1 << bn))
    claims = [evaluation(gate, [q], [], L, Rr) for q in qs]
    return [L, Rr], claims, qs, gate


# -------------------------------------------------------------------- gkr/
class Proof:
    def __init__(self, n_layers):
        self.SumcheckProofs = [[] for _ in range(n_layers)]
        self.Claims = [[] for _ in range(n_layers)]
        self.QPrimes = [[] for _ in range(n_layers)]


def gkr_prove(c, a, qprime):
    """gkr/prover.go:21-91 (a is consumed conceptually; here it is only read)"""
    n_layers = len(c)
    proof = Proof(n_layers)
    proof.QPrimes[n_layers - 1] = [list(qprime)]
    for layer in range(n_layers - 1, -1, -1):
        if not c[layer].In:
            break
        X = [a[p] for p in c[layer].In]  # circuit/assignment.go:35-57
        sp, next_q, final = sumcheck_prove(X, proof.QPrimes[layer], proof.Claims[layer], c[layer].Gate)
        proof.SumcheckProofs[layer] = sp
        for i in range(1, len(final)):
            inp = c[layer].In[i - 1]
            if len(proof.Claims[inp]) < 1:
                proof.Claims[inp] = [0] * len(c[inp].Out)
                proof.QPrimes[inp] = [None] * len(c[inp].Out)
            at = c[inp].Out.index(layer)
            proof.Claims[inp][at] = final[i]
            proof.QPrimes[inp][at] = next_q
    return proof


def gkr_verify(c, proof, inputs, outputs, qprime):
    """gkr/verifier.go:15-132; raises ValueError on rejection"""
    n_layers = len(c)
    if list(qprime) != list(proof.QPrimes[n_layers - 1][0]):
        raise ValueError("initial qPrime mismatch")
    claims = [list(x) for x in proof.Claims]
    claims[n_layers - 1] = claims[n_layers - 1] + [evaluate(outputs, qprime)]
    for layer in range(n_layers - 1, -1, -1):
        if not c[layer].In:
            break
        next_q, next_claim, rho = sumcheck_verify(claims[layer], proof.SumcheckProofs[layer])
        sub = []
        for inp in c[layer].In:
            at = c[inp].Out.index(layer)
            if proof.QPrimes[inp][at] != next_q:
                raise ValueError("qPrime mismatch at layer %d" % layer)
            sub.append(claims[inp][at])
        expected = c[layer].Gate.eval(*sub)
        tmp = [eval_eq(qp, next_q) for qp in proof.QPrimes[layer]]
        expected = expected * eval_univariate(tmp, rho) % Q
        if expected != next_claim:
            raise ValueError("final claim mismatch at layer %d" % layer)
    for layer in range(len(inputs)):
        if evaluate(inputs[layer], proof.QPrimes[layer][0]) != claims[layer][0]:
            raise ValueError("input layer %d check failed" % layer)
    return True


def gkr_proof_to_vec(proof):
    """prover/gadget/hints.go:236-271 (regular form)"""
    vec = []
    for layer in proof.SumcheckProofs:
        for rnd in layer:
            vec.extend(rnd)
    for layer in proof.Claims:
        vec.extend(layer)
    for layer in proof.QPrimes:
        for qs in layer:
            vec.extend(qs)
    return vec


def proof_vec_len(bn):
    """prover/gadget/hints.go:76-116 for the MiMC circuit: 1006*bn + 183"""
    return 1006 * bn + 183


# ------------------------------------------------------------ conversions
def to_mont_limbs(v):
    m = v * R % Q
    return [(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_mont_limbs(l):
    m = l[0] | (l[1] << 64) | (l[2] << 128) | (l[3] << 192)
    return m * pow(R, Q - 2, Q) % Q
