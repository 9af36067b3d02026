"""ctypes binding of the C oracle (oracle/gkr_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  Never imported by the product package.

Field elements cross this binding as numpy uint64 arrays of shape (..., 4): Montgomery form,
little-endian limbs -- byte-identical to Go's []fr.Element.
"""
import ctypes
import os
import subprocess

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_here, "_build", "liboracle.so")
N_LAYERS = 94
GATE_IDENTITY, GATE_CIPHER = 0, 1

Q = 21888242871839275222246405745257275088548364400416034343698204186575808495617
R = (1 << 256) % Q
RINV = pow(R, Q - 2, Q)


def build(force=False):
    srcs = [os.path.join(_here, f) for f in ("gkr_oracle.c", "fr.h", "arks.inc", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _here, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_proof_vec_len.restype = ctypes.c_size_t
        _lib.orc_proof_vec_len.argtypes = [ctypes.c_size_t]
        _lib.orc_get_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _fr(shape=()):
    return np.zeros(tuple(shape) + (4,), dtype=np.uint64)


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.shape[-1] == 4
    return a


# ------------------------------------------------------------ int <-> limbs helpers
def to_mont(vals):
    """list of python ints (regular form) -> (n,4) uint64 Montgomery limbs"""
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        m = (v % Q) * R % Q
        for j in range(4):
            out[i, j] = (m >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def from_mont(arr):
    """(...,4) uint64 Montgomery limbs -> flat list of python ints (regular form)"""
    a = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
    res = []
    for row in a.tolist():
        m = row[0] | (row[1] << 64) | (row[2] << 128) | (row[3] << 192)
        res.append(m * RINV % Q)
    return res


def set_threads(n):
    lib().orc_set_threads(ctypes.c_int(n))


def get_threads():
    return lib().orc_get_threads()


# ------------------------------------------------------------ primitives
def fr_mul(a, b):
    a, b = _c(a), _c(b)
    z = _fr()
    lib().orc_fr_mul(_p(a), _p(b), _p(z))
    return z


def fr_mul_portable(a, b):
    """the unsigned __int128 CIOS form (cross-check of the MULX/ADX multiplier)"""
    a, b = _c(a), _c(b)
    z = _fr()
    lib().orc_fr_mul_portable(_p(a), _p(b), _p(z))
    return z


def fr_mul_kind():
    lib().orc_fr_mul_kind.restype = ctypes.c_char_p
    return lib().orc_fr_mul_kind().decode()


def bench_fr_mul(iters=2_000_000):
    """(ns per independent product, ns per dependent product) on one thread"""
    out = (ctypes.c_double * 2)()
    lib().orc_bench_fr_mul(ctypes.c_size_t(iters), out)
    return float(out[0]), abs(float(out[1]))


def fr_add(a, b):
    a, b = _c(a), _c(b)
    z = _fr()
    lib().orc_fr_add(_p(a), _p(b), _p(z))
    return z


def fr_sub(a, b):
    a, b = _c(a), _c(b)
    z = _fr()
    lib().orc_fr_sub(_p(a), _p(b), _p(z))
    return z


def fr_inv(a):
    a = _c(a)
    z = _fr()
    lib().orc_fr_inv(_p(a), _p(z))
    return z


def mimc_hash(inputs):
    x = _c(inputs).reshape(-1, 4)
    out = _fr()
    lib().orc_mimc_hash(_p(x), ctypes.c_size_t(x.shape[0]), _p(out))
    return out


def mimc_keyed_permutation(x, key):
    out = _fr()
    lib().orc_mimc_keyed_permutation(_p(_c(x)), _p(_c(key)), _p(out))
    return out


def random_fr_array(n):
    out = _fr((n,))
    lib().orc_random_fr_array(_p(out), ctypes.c_size_t(n))
    return out


def fold(tab, r):
    t = _c(tab).copy()
    lib().orc_fold(_p(t), ctypes.c_size_t(t.shape[0]), _p(_c(r)))
    return t[: t.shape[0] // 2].copy()


def evaluate(tab, coords):
    t, c = _c(tab), _c(coords).reshape(-1, 4)
    out = _fr()
    lib().orc_evaluate(_p(t), ctypes.c_size_t(t.shape[0]), _p(c), ctypes.c_size_t(c.shape[0]), _p(out))
    return out


def eval_eq(q, h):
    q, h = _c(q).reshape(-1, 4), _c(h).reshape(-1, 4)
    out = _fr()
    lib().orc_eval_eq(_p(q), _p(h), ctypes.c_size_t(q.shape[0]), _p(out))
    return out


def folded_eq_table(q, multiplier=None):
    q = _c(q).reshape(-1, 4)
    n = q.shape[0]
    out = _fr((1 << n,))
    m = _c(multiplier) if multiplier is not None else None
    lib().orc_folded_eq_table(_p(out), _p(q), ctypes.c_size_t(n), _p(m))
    return out


def chunked_eq_table(q, chunk_size, multiplier=None):
    q = _c(q).reshape(-1, 4)
    n = q.shape[0]
    out = _fr((1 << n,))
    m = _c(multiplier) if multiplier is not None else None
    for cid in range((1 << n) // chunk_size):
        lib().orc_chunk_of_eq_table(_p(out), ctypes.c_size_t(cid), ctypes.c_size_t(chunk_size), _p(q), ctypes.c_size_t(n), _p(m))
    return out


def eval_univariate(coeffs, x):
    c = _c(coeffs).reshape(-1, 4)
    out = _fr()
    lib().orc_eval_univariate(_p(c), ctypes.c_size_t(c.shape[0]), _p(_c(x)), _p(out))
    return out


def lagrange_coefficient(domain):
    out = _fr((domain, domain))
    lib().orc_lagrange_coefficient(ctypes.c_int(domain), _p(out))
    return out


def interpolate_on_range(values):
    v = _c(values).reshape(-1, 4)
    out = _fr((v.shape[0],))
    lib().orc_interpolate_on_range(_p(v), ctypes.c_size_t(v.shape[0]), _p(out))
    return out


def make_eq_table(claims, qprimes):
    """qprimes: (n_q, bn, 4); claims: (n_claims, 4) -> (eq, rho)"""
    qp = _c(qprimes)
    n_q, bn = qp.shape[0], qp.shape[1]
    cl = _c(claims).reshape(-1, 4) if claims is not None and len(claims) else _fr((0,))
    eq, rho = _fr((1 << bn,)), _fr()
    lib().orc_make_eq_table(_p(cl), ctypes.c_size_t(cl.shape[0]), _p(qp), ctypes.c_size_t(n_q), ctypes.c_size_t(bn), _p(eq), _p(rho))
    return eq, rho


def partial_evals(eq, x0, x1, gate_kind, ark=None):
    eq, x0 = _c(eq), _c(x0)
    x1 = _c(x1) if x1 is not None else None
    a = _c(ark) if ark is not None else None
    out = _fr((9 if gate_kind == GATE_CIPHER else 3,))
    lib().orc_partial_evals(_p(eq), _p(x0), _p(x1), ctypes.c_size_t(eq.shape[0]), ctypes.c_int(gate_kind), _p(a), _p(out))
    return out


def sumcheck_prove(X, qprimes, claims, gate_kind, ark=None):
    """X: list of 1 or 2 (2^bn,4) tables (copied, not consumed). Returns (proof (bn,ncoef,4), challenges, final_claims)."""
    qp = _c(qprimes)
    n_q, bn = qp.shape[0], qp.shape[1]
    cl = _c(claims).reshape(-1, 4) if claims is not None and len(claims) else _fr((0,))
    x0 = _c(X[0]).copy()
    x1 = _c(X[1]).copy() if gate_kind == GATE_CIPHER else None
    nco = 9 if gate_kind == GATE_CIPHER else 3
    nin = 2 if gate_kind == GATE_CIPHER else 1
    proof, chal, fin = _fr((bn, nco)), _fr((bn,)), _fr((1 + nin,))
    a = _c(ark) if ark is not None else None
    lib().orc_sumcheck_prove(_p(x0), _p(x1), ctypes.c_size_t(bn), _p(qp), ctypes.c_size_t(n_q), _p(cl), ctypes.c_size_t(cl.shape[0]),
                             ctypes.c_int(gate_kind), _p(a), _p(proof), _p(chal), _p(fin))
    return proof, chal, fin


def sumcheck_verify(claims, proof):
    cl = _c(claims).reshape(-1, 4)
    pr = _c(proof)
    bn, nco = pr.shape[0], pr.shape[1]
    chal, fin, rho = _fr((bn,)), _fr(), _fr()
    rc = lib().orc_sumcheck_verify(_p(cl), ctypes.c_size_t(cl.shape[0]), _p(pr), ctypes.c_size_t(bn), ctypes.c_size_t(nco), _p(chal), _p(fin), _p(rho))
    return rc, chal, fin, rho


def evaluation(gate_kind, ark, qprimes, claims, x0, x1=None):
    qp = _c(qprimes)
    n_q, bn = qp.shape[0], qp.shape[1]
    cl = _c(claims).reshape(-1, 4) if claims is not None and len(claims) else _fr((0,))
    a = _c(ark) if ark is not None else None
    out = _fr()
    lib().orc_evaluation(ctypes.c_int(gate_kind), _p(a), _p(qp), ctypes.c_size_t(n_q), ctypes.c_size_t(bn), _p(cl), ctypes.c_size_t(cl.shape[0]),
                         _p(_c(x0)), _p(_c(x1)) if x1 is not None else None, _p(out))
    return out


def mimc_assign(key, msg):
    """-> (94, n, 4) array of all layers (circuit/assignment.go:12-32)"""
    key, msg = _c(key), _c(msg)
    n = key.shape[0]
    layers = _fr((N_LAYERS, n))
    ptrs = (ctypes.c_void_p * N_LAYERS)(*[layers[l].ctypes.data for l in range(N_LAYERS)])
    lib().orc_mimc_assign(_p(key), _p(msg), ctypes.c_size_t(n), ptrs)
    return layers


def gkr_prove_mimc(layers, qprime):
    """layers: (94,n,4) (copied; the C call consumes its copy). -> flat proof vector (1006*bn+183, 4), Montgomery."""
    L = _c(layers).copy()
    n = L.shape[1]
    bn = n.bit_length() - 1
    qp = _c(qprime).reshape(-1, 4) if bn else _fr((0,))
    vec = _fr((proof_vec_len(bn),))
    ptrs = (ctypes.c_void_p * N_LAYERS)(*[L[l].ctypes.data for l in range(N_LAYERS)])
    lib().orc_gkr_prove_mimc(ptrs, ctypes.c_size_t(bn), _p(qp), _p(vec))
    return vec


def gkr_verify_mimc(vec, in0, in1, outputs, qprime):
    in0 = _c(in0)
    bn = in0.shape[0].bit_length() - 1
    qp = _c(qprime).reshape(-1, 4) if bn else _fr((0,))
    return lib().orc_gkr_verify_mimc(_p(_c(vec)), ctypes.c_size_t(bn), _p(in0), _p(_c(in1)), _p(_c(outputs)), _p(qp))


def assign_and_prove_mimc(key, msg, qprime):
    """Reference flow Assign + Prove on the CPU. -> (out93 (n,4), vec)"""
    key, msg = _c(key), _c(msg)
    n = key.shape[0]
    bn = n.bit_length() - 1
    qp = _c(qprime).reshape(-1, 4) if bn else _fr((0,))
    out93, vec = _fr((n,)), _fr((proof_vec_len(bn),))
    rc = lib().orc_assign_and_prove_mimc(_p(key), _p(msg), ctypes.c_size_t(bn), _p(qp), _p(out93), _p(vec))
    if rc:
        raise MemoryError("oracle: allocation failed")
    return out93, vec


def proof_vec_len(bn):
    return int(lib().orc_proof_vec_len(ctypes.c_size_t(bn)))
