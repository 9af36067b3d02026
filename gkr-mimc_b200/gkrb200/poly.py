"""poly/ of the reference on the device."""
from ._lib import check, lib
from .context import _p, fr_array, fr_empty


def FoldedEqTable(ctx, qPrime, multiplier=None):
    """poly/eq.go:41-59 (also what ChunkOfEqTable :62-89 assembles)"""
    q = fr_array(qPrime).reshape(-1, 4)
    bn = q.shape[0]
    out = fr_empty(1 << bn)
    m = fr_array(multiplier).reshape(1, 4) if multiplier is not None else None
    check(lib().gkrb200_eq_table(ctx.handle, _p(q), 1, bn, _p(m), _p(out)))
    return out


def MultiEqTable(ctx, qPrimes, multipliers):
    """sum_j multipliers[j] * eq(qPrimes[j], .)  -- the combination of sumcheck/prover.go:121-141"""
    q = fr_array(qPrimes)
    n_q, bn = q.shape[0], q.shape[1]
    m = fr_array(multipliers).reshape(n_q, 4)
    out = fr_empty(1 << bn)
    check(lib().gkrb200_eq_table(ctx.handle, _p(q), n_q, bn, _p(m), _p(out)))
    return out


def Fold(ctx, table, r):
    """MultiLin.Fold (poly/multilin.go:19-36); returns the folded table (half the length)"""
    t = fr_array(table).reshape(-1, 4)
    out = fr_empty(t.shape[0] // 2)
    check(lib().gkrb200_fold(ctx.handle, _p(t), t.shape[0], _p(fr_array(r)), _p(out)))
    return out


def InterpolateOnRange(values):
    """poly/lagrange.go:96-111"""
    v = fr_array(values).reshape(-1, 4)
    out = fr_empty(v.shape[0])
    check(lib().gkrb200_interpolate(_p(v), v.shape[0], _p(out)))
    return out


def Evaluate(ctx, table, coords):
    """MultiLin.Evaluate (poly/multilin.go:59-66) on the device"""
    t = fr_array(table).reshape(-1, 4)
    q = fr_array(coords).reshape(-1, 4) if t.shape[0] > 1 else None
    out = fr_empty(1)
    check(lib().gkrb200_mle_evaluate(ctx.handle, _p(t), t.shape[0], _p(q), _p(out)))
    return out[0]


def Convert(ctx, values, to_montgomery):
    """batched fr.Element.SetBigInt / ToBigIntRegular on the device"""
    v = fr_array(values).reshape(-1, 4)
    out = fr_empty(v.shape[0])
    check(lib().gkrb200_convert(ctx.handle, _p(v), v.shape[0], _p(out), 1 if to_montgomery else 0))
    return out


def EvalUnivariate(coeffs, x):
    """poly/lagrange.go:31-39 (host; coefficients low -> high)"""
    c = fr_array(coeffs).reshape(-1, 4)
    out = fr_empty()
    check(lib().gkrb200_eval_univariate(_p(c), c.shape[0], _p(fr_array(x)), _p(out)))
    return out


def EvalEq(qPrime, nextQPrime):
    """poly/eq.go:19-32 (host)"""
    q, h = fr_array(qPrime).reshape(-1, 4), fr_array(nextQPrime).reshape(-1, 4)
    if q.shape != h.shape:
        raise ValueError("EvalEq: the two points have different sizes")
    out = fr_empty()
    check(lib().gkrb200_eval_eq(_p(q) if q.shape[0] else None, _p(h) if q.shape[0] else None, q.shape[0], _p(out)))
    return out
