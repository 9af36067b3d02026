"""ctypes binding of the C oracle for the FFT half of the Groth16 prover (oracle/fft_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/ and by tools' cpu_baseline legs.  Never imported by the product package.
Field elements cross as numpy uint64 arrays (..., 4) = Go's []fr.Element (Montgomery, little-endian limbs).
"""
import ctypes
import os
import subprocess

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_here, "_build", "libfftoracle.so")
DIT, DIF = 0, 1


def build(force=False):
    srcs = [os.path.join(_here, f) for f in ("fft_oracle.c", "fr.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _here, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    return a


def next_pow2(m):
    n = 1
    while n < m:
        n <<= 1
    return n


def fft(a, decimation, coset=0):
    a = _c(a).copy()
    assert a.shape[0] == next_pow2(a.shape[0])
    lib().orc_fft(_p(a), ctypes.c_size_t(a.shape[0]), ctypes.c_int(decimation), ctypes.c_int(coset))
    return a


def fft_inverse(a, decimation, coset=0):
    a = _c(a).copy()
    assert a.shape[0] == next_pow2(a.shape[0])
    lib().orc_fft_inverse(_p(a), ctypes.c_size_t(a.shape[0]), ctypes.c_int(decimation), ctypes.c_int(coset))
    return a


def compute_h(a, b, c, cardinality=None):
    """prover/gadget/prove.go:310-366 -> (cardinality, 4) REGULAR-form words, bit-reversed coefficient order"""
    a, b, c = _c(a), _c(b), _c(c)
    assert a.shape == b.shape == c.shape
    n = cardinality or next_pow2(a.shape[0])
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().orc_compute_h(_p(a), _p(b), _p(c), ctypes.c_size_t(a.shape[0]), ctypes.c_size_t(n), _p(out))
    return out


def domain(cardinality):
    """-> (Generator, FinerGenerator, CardinalityInv), each (4,) Montgomery"""
    out = np.zeros(12, dtype=np.uint64)
    lib().orc_fft_domain(ctypes.c_size_t(cardinality), _p(out))
    return out[:4].copy(), out[4:8].copy(), out[8:].copy()


def mul_elementwise(a, b):
    a, b = _c(a), _c(b)
    out = np.zeros_like(a)
    lib().orc_fr_mul_elementwise(_p(a), _p(b), ctypes.c_size_t(a.shape[0]), _p(out))
    return out


def eval_poly_bitrev(coef, z, regular=True):
    """P(z), P's coefficient rev(i) at coef[i] (computeH's output order); z, result: (4,) Montgomery"""
    coef, z = _c(coef), np.ascontiguousarray(z, dtype=np.uint64).reshape(4)
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_eval_poly_bitrev(_p(coef), ctypes.c_size_t(coef.shape[0]), ctypes.c_int(1 if regular else 0), _p(z), _p(out))
    return out


def eval_lagrange(a, n, z):
    """A(z) for deg A < n, A(w^i) = a[i] for i < len(a), 0 on the rest of the domain; z outside the domain; (4,) Montgomery"""
    a, z = _c(a), np.ascontiguousarray(z, dtype=np.uint64).reshape(4)
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_eval_lagrange(_p(a), ctypes.c_size_t(a.shape[0]), ctypes.c_size_t(n), _p(z), _p(out))
    return out


def quotient_identity_holds(a, b, c, h_regular, n, z):
    """H(z) (z^n - 1) == A(z) B(z) - C(z), everything evaluated by this oracle at the Montgomery point z (Python integers inside)"""
    Q = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    rinv = pow((1 << 256) % Q, -1, Q)
    val = lambda limbs: sum(int(x) << (64 * j) for j, x in enumerate(limbs)) * rinv % Q
    zz = val(np.ascontiguousarray(z, dtype=np.uint64).reshape(4))
    A, B, C = (val(eval_lagrange(v, n, z)) for v in (a, b, c))
    H = val(eval_poly_bitrev(h_regular, z, regular=True))
    return H * (pow(zz, n, Q) - 1) % Q == (A * B - C) % Q
