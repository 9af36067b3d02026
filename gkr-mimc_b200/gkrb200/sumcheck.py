"""sumcheck/ of the reference (prover side) on the device."""
from ._lib import check, lib
from .context import _p, fr_array, fr_empty
from .gates import GATE_CIPHER


def Prove(ctx, X, qPrimes, claims, gate):
    """sumcheck.Prove (sumcheck/prover.go:46-90) -> (proof[bn][deg+2], challenges[bn], finalClaims)

    X: list of tables (1 for IdentityGate, 2 for CipherGate); qPrimes: (n_q, bn, 4); claims: (n_claims, 4).
    Unlike the reference the caller's tables are left untouched (the device works on its own copy)."""
    q = fr_array(qPrimes)
    n_q, bn = q.shape[0], q.shape[1]
    cl = fr_array(claims).reshape(-1, 4) if claims is not None and len(claims) else None
    n_claims = 0 if cl is None else cl.shape[0]
    for i, x in enumerate(X):
        if len(x) != 1 << bn:
            raise ValueError("inconsistent sizes : bn is %d but table %d has size %d" % (bn, i, len(x)))  # prover.go:54
    x0 = fr_array(X[0]).reshape(-1, 4)
    x1 = fr_array(X[1]).reshape(-1, 4) if gate.kind == GATE_CIPHER else None
    nco = gate.Degree() + 2
    nin = 2 if gate.kind == GATE_CIPHER else 1
    proof, chal, fin = fr_empty(bn, nco), fr_empty(bn), fr_empty(1 + nin)
    ark = fr_array(gate.ark) if gate.ark is not None else None
    check(lib().gkrb200_sumcheck_prove(ctx.handle, _p(x0), _p(x1), bn, _p(q), n_q, _p(cl), n_claims, gate.kind, _p(ark),
                                       _p(proof), _p(chal), _p(fin)))
    return proof, chal, fin


def ProveDevice(ctx, d_X, qPrimes, claims, gate):
    """sumcheck.Prove with the tables already resident on the device: d_X = raw device pointers (Go layout, 2^bn entries each)"""
    import ctypes
    q = fr_array(qPrimes)
    n_q, bn = q.shape[0], q.shape[1]
    cl = fr_array(claims).reshape(-1, 4) if claims is not None and len(claims) else None
    n_claims = 0 if cl is None else cl.shape[0]
    nco = gate.Degree() + 2
    nin = 2 if gate.kind == GATE_CIPHER else 1
    proof, chal, fin = fr_empty(bn, nco), fr_empty(bn), fr_empty(1 + nin)
    ark = fr_array(gate.ark) if gate.ark is not None else None
    x1 = ctypes.c_void_p(d_X[1]) if gate.kind == GATE_CIPHER else None
    check(lib().gkrb200_sumcheck_prove_device(ctx.handle, ctypes.c_void_p(d_X[0]), x1, bn, _p(q), n_q, _p(cl), n_claims, gate.kind, _p(ark),
                                              _p(proof), _p(chal), _p(fin)))
    return proof, chal, fin


def PartialEvals(ctx, eq, X, gate):
    """one call of getPartialPolyChunk over the whole table (sumcheck/algo.go:54-205)"""
    e = fr_array(eq).reshape(-1, 4)
    x0 = fr_array(X[0]).reshape(-1, 4)
    x1 = fr_array(X[1]).reshape(-1, 4) if gate.kind == GATE_CIPHER else None
    out = fr_empty(gate.Degree() + 2)
    ark = fr_array(gate.ark) if gate.ark is not None else None
    check(lib().gkrb200_round_eval(ctx.handle, _p(e), _p(x0), _p(x1), e.shape[0], gate.kind, _p(ark), _p(out)))
    return out


def Verify(claims, proof):
    """sumcheck.Verify (sumcheck/verifier.go:28-65) -> (challenges, finalClaim, recombChal); raises GkrB200Error
    (code GKRB200_ERR_VERIFY) where the reference returns an error.  Host-side, like the reference's verifier."""
    cl = fr_array(claims).reshape(-1, 4)
    pr = fr_array(proof)
    bn = pr.shape[0] if pr.ndim == 3 else 0
    nco = pr.shape[1] if bn else 1
    chal, fin, rho = fr_empty(bn), fr_empty(), fr_empty()
    check(lib().gkrb200_sumcheck_verify(_p(cl), cl.shape[0], _p(pr) if bn else None, bn, nco, _p(chal) if bn else None, _p(fin), _p(rho)))
    return chal, fin, rho


def Evaluation(ctx, gate, qPrimes, claims, *X):
    """sumcheck.Evaluation (sumcheck/instance.go:49-68): sum_x Eq(x) * gate(X(x)) with Eq = sum_j rho^j eq(qPrimes[j], .),
    rho = GetChallenge(claims) when there are several qPrimes.  The gate values are formed on the device and
    sum_x eq(q, x) g(x) is MultiLin.Evaluate of g at q, also on the device."""
    import numpy as np
    from .common import GetChallenge
    q = fr_array(qPrimes)
    n_q, bn = q.shape[0], q.shape[1]
    x0 = fr_array(X[0]).reshape(-1, 4)
    if gate.kind == GATE_CIPHER:
        t = ctx.fr_batch(1, x0, fr_array(X[1]).reshape(-1, 4))
        t = ctx.fr_batch(1, t, np.ascontiguousarray(np.broadcast_to(gate.ark, t.shape)))
        g = ctx.fr_batch(3, t)
    else:
        g = x0
    L = lib()
    res, pw = None, None
    if claims is None or len(claims) < 1:
        n_q = 1  # prover.go:117-119: without claims only the first qPrime is used
    rho = GetChallenge(claims) if n_q > 1 else None
    for j in range(n_q):
        e = fr_empty()
        check(L.gkrb200_mle_evaluate(ctx.handle, _p(g), g.shape[0], _p(np.ascontiguousarray(q[j])) if bn else None, _p(e)))
        if j == 0:
            res = e
            continue
        pw = rho.copy() if pw is None else _scalar(0, pw, rho)
        res = _scalar(1, res, _scalar(0, e, pw))
    return res


def _scalar(op, a, b):
    out = fr_empty()
    check(lib().gkrb200_fr_scalar(op, _p(fr_array(a)), _p(fr_array(b)), _p(out)))
    return out
