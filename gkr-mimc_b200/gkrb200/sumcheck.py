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


def PartialEvals(ctx, eq, X, gate):
    """one call of getPartialPolyChunk over the whole table (sumcheck/algo.go:54-205)"""
    e = fr_array(eq).reshape(-1, 4)
    x0 = fr_array(X[0]).reshape(-1, 4)
    x1 = fr_array(X[1]).reshape(-1, 4) if gate.kind == GATE_CIPHER else None
    out = fr_empty(gate.Degree() + 2)
    ark = fr_array(gate.ark) if gate.ark is not None else None
    check(lib().gkrb200_round_eval(ctx.handle, _p(e), _p(x0), _p(x1), e.shape[0], gate.kind, _p(ark), _p(out)))
    return out
