"""common/ of the reference: the Fiat-Shamir challenge (host, serial by nature)."""
import numpy as np

from ._lib import check, lib
from .context import _p, fr_array, fr_empty


def GetChallenge(seed):
    """common/challenge.go:10 -> hash.MimcHash (hash/mimc.go:11-18)"""
    s = fr_array(seed).reshape(-1, 4)
    out = fr_empty()
    check(lib().gkrb200_mimc_hash(_p(s), s.shape[0], _p(out)))
    return out


MimcHash = GetChallenge


def ToMontgomery(regular):
    a = fr_array(regular)
    out = np.empty_like(a)
    check(lib().gkrb200_to_montgomery(_p(a), a.size // 4, _p(out)))
    return out


def FromMontgomery(mont):
    a = fr_array(mont)
    out = np.empty_like(a)
    check(lib().gkrb200_from_montgomery(_p(a), a.size // 4, _p(out)))
    return out


def SetUint64(values):
    """fr.Element.SetUint64 for an iterable of python ints < 2^64"""
    vals = list(values)
    a = np.zeros((len(vals), 4), dtype=np.uint64)
    if vals:
        a[:, 0] = np.array(vals, dtype=np.uint64)
    return ToMontgomery(a)


def RandomFrArray(n):
    """common/common.go:49-55: res[i] = SetUint64(i*i XOR 0xf45c9df123f) (wrapping uint64 multiply)"""
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        v = (i * i) ^ np.uint64(0xF45C9DF123F)
    a = np.zeros((n, 4), dtype=np.uint64)
    a[:, 0] = v
    return ToMontgomery(a)


def Ark(i):
    """hash.Arks[i] (hash/ark.go:232-336), the constant of the cipher gate of layer i+3 (examples/mimc.go:29)"""
    out = fr_empty()
    check(lib().gkrb200_mimc_ark(i, _p(out)))
    return out
