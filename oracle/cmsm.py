"""ctypes binding of the C oracle for the G1 multi-exponentiation path (oracle/msm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/ and __graft_entry__.smoke().  Never imported by the product package.

Points cross this binding as numpy uint64 arrays of shape (..., 8) = Go's []bn254.G1Affine (X then Y, 4 little-endian limbs each,
Montgomery form, infinity = all zero); scalars as (..., 4) = []fr.Element, Montgomery or regular form as stated per call.
"""
import ctypes
import os
import subprocess

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_here, "_build", "libmsmoracle.so")

P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
Q = 21888242871839275222246405745257275088548364400416034343698204186575808495617
RP = (1 << 256) % P
RP_INV = pow(RP, -1, P)
RQ = (1 << 256) % Q


def build(force=False):
    srcs = [os.path.join(_here, f) for f in ("msm_oracle.c", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _here, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_g1_is_on_curve.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, w):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.shape[-1] == w
    return a


def threads():
    return max(1, min(32, os.cpu_count() or 1))


# ------------------------------------------------------------ python int <-> limbs
def limbs(v, n=4):
    return [(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(n)]


def unlimbs(row):
    return sum(int(x) << (64 * j) for j, x in enumerate(row))


def scalars_regular(vals):
    """python ints -> (n, 4) regular-form fr limbs"""
    return np.array([limbs(v % Q) for v in vals], dtype=np.uint64).reshape(-1, 4)


def scalars_mont(vals):
    return np.array([limbs((v % Q) * RQ % Q) for v in vals], dtype=np.uint64).reshape(-1, 4)


def point_to_ints(pt):
    """(8,) Montgomery limbs -> (x, y) python ints, (0, 0) for infinity"""
    pt = np.asarray(pt, dtype=np.uint64).reshape(8)
    return (unlimbs(pt[:4]) * RP_INV % P, unlimbs(pt[4:]) * RP_INV % P)


def point_from_ints(xy):
    x, y = xy
    return np.array(limbs(x * RP % P) + limbs(y * RP % P), dtype=np.uint64)


# ------------------------------------------------------------ calls
def generator():
    out = np.zeros(8, dtype=np.uint64)
    lib().orc_g1_generator(_p(out))
    return out


def is_on_curve(pt):
    pt = _c(pt, 8)
    return bool(lib().orc_g1_is_on_curve(_p(pt)))


def add(a, b):
    a, b = _c(a, 8), _c(b, 8)
    out = np.zeros(8, dtype=np.uint64)
    lib().orc_g1_add(_p(a), _p(b), _p(out))
    return out


def neg(a):
    a = _c(a, 8)
    out = np.zeros(8, dtype=np.uint64)
    lib().orc_g1_neg(_p(a), _p(out))
    return out


def scalar_mul(pt, k):
    """k: python int (regular value)"""
    pt = _c(pt, 8)
    kk = np.array(limbs(k % Q), dtype=np.uint64)
    out = np.zeros(8, dtype=np.uint64)
    lib().orc_g1_scalar_mul(_p(pt), _p(kk), _p(out))
    return out


def multiexp(points, scalars, mont=False, nthreads=None):
    points, scalars = _c(points, 8).reshape(-1, 8), _c(scalars, 4).reshape(-1, 4)
    assert points.shape[0] == scalars.shape[0]
    out = np.zeros(8, dtype=np.uint64)
    lib().orc_g1_multiexp(_p(points), _p(scalars), ctypes.c_size_t(points.shape[0]), ctypes.c_int(1 if mont else 0),
                          ctypes.c_int(nthreads or threads()), _p(out))
    return out


def multiexp_buckets(points, scalars, mont=False, nthreads=None):
    """G1Affine.MultiExp by gnark-crypto's published algorithm (bucket method, one task per window): the CPU baseline and a fast second oracle"""
    points, scalars = _c(points, 8).reshape(-1, 8), _c(scalars, 4).reshape(-1, 4)
    assert points.shape[0] == scalars.shape[0]
    out = np.zeros(8, dtype=np.uint64)
    lib().orc_g1_multiexp_buckets(_p(points), _p(scalars), ctypes.c_size_t(points.shape[0]), ctypes.c_int(1 if mont else 0),
                                  ctypes.c_int(nthreads or threads()), _p(out))
    return out


def gen_points(n, a=0x1234567, b=0x9E3779B97F4A7C15):
    """P_i = (a + i*b) * G, i < n"""
    out = np.zeros((n, 8), dtype=np.uint64)
    aa, bb = np.array(limbs(a % Q), dtype=np.uint64), np.array(limbs(b % Q), dtype=np.uint64)
    if n:
        lib().orc_g1_gen_points(ctypes.c_size_t(n), _p(aa), _p(bb), _p(out))
    return out


def keccak256(data: bytes) -> bytes:
    buf = (ctypes.c_uint8 * max(1, len(data))).from_buffer_copy(data or b"\0")
    out = (ctypes.c_uint8 * 32)()
    lib().orc_keccak256(buf, ctypes.c_size_t(len(data)), out)
    return bytes(out)


def raw_bytes(pt) -> bytes:
    pt = _c(pt, 8)
    out = (ctypes.c_uint8 * 64)()
    lib().orc_g1_raw_bytes(_p(pt), out)
    return bytes(out)


def derive_randomness_from_point(pt):
    """-> (4,) regular-form fr limbs"""
    pt = _c(pt, 8)
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_derive_randomness_from_point(_p(pt), _p(out))
    return out


def initial_randomness(pub_points, pub_scalars, priv_points, priv_scalars, mont=False):
    """hints.go:162-192 -> (KrsGkrPriv (8,), initialRandomness (4,) regular form)"""
    pp, ps = _c(pub_points, 8).reshape(-1, 8), _c(pub_scalars, 4).reshape(-1, 4)
    qp, qs = _c(priv_points, 8).reshape(-1, 8), _c(priv_scalars, 4).reshape(-1, 4)
    krs_priv = np.zeros(8, dtype=np.uint64)
    rnd = np.zeros(4, dtype=np.uint64)
    lib().orc_initial_randomness(_p(pp), _p(ps), ctypes.c_size_t(pp.shape[0]), _p(qp), _p(qs), ctypes.c_size_t(qp.shape[0]),
                                 ctypes.c_int(1 if mont else 0), ctypes.c_int(threads()), _p(krs_priv), _p(rnd))
    return krs_priv, rnd


# ------------------------------------------------------------ G2 (points: (..., 16) uint64 = []bn254.G2Affine: X.A0, X.A1, Y.A0, Y.A1)
def g2_generator():
    out = np.zeros(16, dtype=np.uint64)
    lib().orc_g2_generator(_p(out))
    return out


def g2_is_on_curve(pt):
    pt = _c(pt, 16)
    lib().orc_g2_is_on_curve.restype = ctypes.c_int
    return bool(lib().orc_g2_is_on_curve(_p(pt)))


def g2_add(a, b):
    a, b = _c(a, 16), _c(b, 16)
    out = np.zeros(16, dtype=np.uint64)
    lib().orc_g2_add(_p(a), _p(b), _p(out))
    return out


def g2_neg(a):
    a = _c(a, 16)
    out = np.zeros(16, dtype=np.uint64)
    lib().orc_g2_neg(_p(a), _p(out))
    return out


def g2_scalar_mul(pt, k):
    pt = _c(pt, 16)
    kk = np.array(limbs(k % Q), dtype=np.uint64)
    out = np.zeros(16, dtype=np.uint64)
    lib().orc_g2_scalar_mul(_p(pt), _p(kk), _p(out))
    return out


def g2_multiexp(points, scalars, mont=False, nthreads=None):
    points, scalars = _c(points, 16).reshape(-1, 16), _c(scalars, 4).reshape(-1, 4)
    assert points.shape[0] == scalars.shape[0]
    out = np.zeros(16, dtype=np.uint64)
    lib().orc_g2_multiexp(_p(points), _p(scalars), ctypes.c_size_t(points.shape[0]), ctypes.c_int(1 if mont else 0),
                          ctypes.c_int(nthreads or threads()), _p(out))
    return out


def g2_gen_points(n, a=0x7654321, b=0xD1B54A32D192ED03):
    """P_i = (a + i*b) * G2gen, i < n"""
    out = np.zeros((n, 16), dtype=np.uint64)
    aa, bb = np.array(limbs(a % Q), dtype=np.uint64), np.array(limbs(b % Q), dtype=np.uint64)
    if n:
        lib().orc_g2_gen_points(ctypes.c_size_t(n), _p(aa), _p(bb), _p(out))
    return out


def g2_point_to_ints(pt):
    """(16,) Montgomery limbs -> ((x0, x1), (y0, y1)) python ints"""
    pt = np.asarray(pt, dtype=np.uint64).reshape(16)
    v = [unlimbs(pt[4 * k:4 * k + 4]) * RP_INV % P for k in range(4)]
    return ((v[0], v[1]), (v[2], v[3]))
