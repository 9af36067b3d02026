"""Host-side mirror of the Groth16-side operations of the gkr-mimc prover over libgkrb200ec.so.

    reference (Go)                                                        here
    --------------------------------------------------------------------  ---------------------------------------------
    pk.pubKGkr / pk.privKGkrSigma []bn254.G1Affine (setup.go:32,116-121)   EcContext.SetBases(slot, points)
    G1Affine.MultiExp(points, scalars, ecc.MultiExpConfig{})               EcContext.MultiExp(slot, scalars) / MultiExpPoints
    G1Affine.Add (hints.go:184)                                            EcContext.Add(a, b)
    G2Affine.MultiExp(pk.G2.B, wireValuesB, cfg) (prove.go:277)            EcContext.SetBasesG2 / MultiExpG2 / MultiExpPointsG2 / AddG2
    DeriveRandomnessFromPoint(g1) (hints.go:147-159)                       DeriveRandomnessFromPoint(g1)
    InitialRandomnessHint.Call (hints.go:162-192)                          EcContext.InitialRandomnessHint(...)
    fft.NewDomain(m, 1, true) (groth16/setup.go:98)                        EcContext.NewDomain(m)
    domain.FFT(a, fft.DIT|DIF, coset) / domain.FFTInverse                  EcContext.FFT(a, decimation, coset) / FFTInverse
    computeH(a, b, c, &pk.Domain) (prove.go:310-366)                       EcContext.ComputeH(a, b, c) / ComputeHDevice
    ComputeGroth16Proof(r1cs, pk, a, b, c, wireValues) (prove.go:100-306)  EcContext.ComputeGroth16Proof(...) (r, s from the caller)

Points are numpy uint64 arrays (..., 8) = []bn254.G1Affine (X, Y; Montgomery; infinity all zero), scalars (..., 4) = []fr.Element.
Everything that touches a point array runs on the GPU; the library fails loudly without one.  There is no CPU fallback.
"""
import ctypes
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_PKG)
SO_PATH = os.path.join(ROOT, "libgkrb200ec.so")

SCALARS_REGULAR, SCALARS_MONTGOMERY = 0, 1
DIT, DIF = 0, 1  # fft.Decimation
MAX_SLOTS = 16


class EcStats(ctypes.Structure):
    _fields_ = [
        ("launches_total", ctypes.c_uint64),
        ("msm_calls", ctypes.c_uint64),
        ("last_n", ctypes.c_uint32),
        ("last_c", ctypes.c_uint32),
        ("last_windows", ctypes.c_uint32),
        ("last_task_size", ctypes.c_uint32),
        ("last_tasks_max", ctypes.c_uint64),
        ("workspace_bytes", ctypes.c_uint64),
        ("h2d_bytes", ctypes.c_uint64),
        ("d2h_bytes", ctypes.c_uint64),
        ("last_device_ms", ctypes.c_double),
        ("fft_calls", ctypes.c_uint64),
        ("last_fft_device_ms", ctypes.c_double),
    ]


class Groth16Pk(ctypes.Structure):
    """gkrb200ec_groth16_pk: which slots hold pk.G1.A / pk.G1.B / pk.G1.Z / pk.G2.B, and the five single points of the key"""
    _fields_ = [("slot_g1_a", ctypes.c_int), ("slot_g1_b", ctypes.c_int), ("slot_g1_z", ctypes.c_int), ("slot_g2_b", ctypes.c_int),
                ("g1_alpha", ctypes.c_void_p), ("g1_beta", ctypes.c_void_p), ("g1_delta", ctypes.c_void_p),
                ("g2_beta", ctypes.c_void_p), ("g2_delta", ctypes.c_void_p)]


class GkrB200EcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("gkrb200ec error %d: %s" % (code, msg))
        self.code = code


_lib = None


def _bind(L):
    """argument types of every entry point of include/gkrb200_ec.h"""
    vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    L.gkrb200ec_version.restype = ctypes.c_char_p
    L.gkrb200ec_last_error.restype = ctypes.c_char_p
    L.gkrb200ec_init.argtypes = [ctypes.POINTER(vp), i32, vp]
    L.gkrb200ec_free.argtypes = [vp]
    L.gkrb200ec_free.restype = None
    L.gkrb200ec_g1_set_bases.argtypes = [vp, i32, vp, sz]
    L.gkrb200ec_g1_multiexp.argtypes = [vp, i32, vp, sz, i32, vp]
    L.gkrb200ec_g1_multiexp_device.argtypes = [vp, i32, vp, sz, i32, vp]
    L.gkrb200ec_g1_multiexp_points.argtypes = [vp, vp, vp, sz, i32, vp]
    L.gkrb200ec_initial_randomness.argtypes = [vp, i32, vp, sz, i32, vp, sz, i32, vp, vp]
    L.gkrb200ec_g1_add.argtypes = [vp, vp, vp, vp]
    L.gkrb200ec_g2_set_bases.argtypes = [vp, i32, vp, sz]
    L.gkrb200ec_g2_multiexp.argtypes = [vp, i32, vp, sz, i32, vp]
    L.gkrb200ec_g2_multiexp_device.argtypes = [vp, i32, vp, sz, i32, vp]
    L.gkrb200ec_g2_multiexp_points.argtypes = [vp, vp, vp, sz, i32, vp]
    L.gkrb200ec_g2_add.argtypes = [vp, vp, vp, vp]
    L.gkrb200ec_g1_raw_bytes.argtypes = [vp, vp]
    L.gkrb200ec_keccak256.argtypes = [vp, sz, vp]
    L.gkrb200ec_derive_randomness_from_point.argtypes = [vp, vp]
    L.gkrb200ec_set_plan.argtypes = [vp, i32, i32]
    L.gkrb200ec_fft_domain_init.argtypes = [vp, ctypes.c_uint64]
    L.gkrb200ec_fft_domain_cardinality.argtypes = [vp]
    L.gkrb200ec_fft_domain_cardinality.restype = ctypes.c_uint64
    L.gkrb200ec_fft.argtypes = [vp, vp, sz, i32, i32]
    L.gkrb200ec_fft_inverse.argtypes = [vp, vp, sz, i32, i32]
    L.gkrb200ec_compute_h.argtypes = [vp, vp, vp, vp, sz, vp, ctypes.POINTER(vp)]
    L.gkrb200ec_groth16_prove.argtypes = [vp, ctypes.POINTER(Groth16Pk), vp, vp, vp, sz, vp, sz, vp, sz, i32, vp, vp, vp, vp, vp]
    L.gkrb200ec_get_stats.argtypes = [vp, ctypes.POINTER(EcStats)]
    return L


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError("gkrb200: %s is missing. Build it with `make -C %s` (or __graft_entry__.build()); there is no CPU fallback."
                           % (SO_PATH, ROOT))
    _lib = _bind(ctypes.CDLL(SO_PATH))
    return _lib


def check(rc):
    if rc != 0:
        raise GkrB200EcError(rc, lib().gkrb200ec_last_error().decode("utf-8", "replace"))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def g1_array(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.shape[-1] != 8:
        raise ValueError("G1Affine points must have a trailing dimension of 8 uint64 limbs")
    return a


def g2_array(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.shape[-1] != 16:
        raise ValueError("G2Affine points must have a trailing dimension of 16 uint64 limbs (X.A0, X.A1, Y.A0, Y.A1)")
    return a


def fr_array(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.shape[-1] != 4:
        raise ValueError("field elements must have a trailing dimension of 4 uint64 limbs")
    return a


# ---- host-only pieces (no device) -------------------------------------------------------------------------------------------
def RawBytes(g1) -> bytes:
    """G1Affine.RawBytes"""
    g1 = g1_array(g1).reshape(8)
    out = (ctypes.c_uint8 * 64)()
    check(lib().gkrb200ec_g1_raw_bytes(_p(g1), out))
    return bytes(out)


def LegacyKeccak256(data: bytes) -> bytes:
    """sha3.NewLegacyKeccak256().Write(data).Sum(nil)"""
    buf = (ctypes.c_uint8 * max(1, len(data))).from_buffer_copy(data if data else b"\0")
    out = (ctypes.c_uint8 * 32)()
    check(lib().gkrb200ec_keccak256(buf, len(data), out))
    return bytes(out)


def DeriveRandomnessFromPoint(g1):
    """prover/gadget/hints.go:147-159 -> (4,) uint64, REGULAR form (the big.Int the hint returns)"""
    g1 = g1_array(g1).reshape(8)
    out = np.zeros(4, dtype=np.uint64)
    check(lib().gkrb200ec_derive_randomness_from_point(_p(g1), _p(out)))
    return out


# ---- device context -----------------------------------------------------------------------------------------------------------
class EcContext:
    def __init__(self, device=0, stream=None):
        self._h = ctypes.c_void_p()
        check(lib().gkrb200ec_init(ctypes.byref(self._h), device, ctypes.c_void_p(stream) if stream else None))
        self.device = device

    def close(self):
        if self._h:
            lib().gkrb200ec_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def SetBases(self, slot, points):
        """upload proving-key points once (pk.pubKGkr, pk.privKGkrSigma, pk.privKNotGkr, pk.G1.A ...)"""
        points = g1_array(points).reshape(-1, 8)
        check(lib().gkrb200ec_g1_set_bases(self._h, slot, _p(points) if points.shape[0] else None, points.shape[0]))

    def MultiExp(self, slot, scalars, form=SCALARS_REGULAR):
        """res.MultiExp(bases[slot][:len(scalars)], scalars, ecc.MultiExpConfig{}) -> (8,) G1Affine"""
        scalars = fr_array(scalars).reshape(-1, 4)
        out = np.zeros(8, dtype=np.uint64)
        check(lib().gkrb200ec_g1_multiexp(self._h, slot, _p(scalars) if scalars.shape[0] else None, scalars.shape[0], form, _p(out)))
        return out

    def MultiExpDevice(self, slot, d_scalars_ptr, n, form=SCALARS_REGULAR):
        """scalars already in device memory (raw pointer, 16-byte aligned, n x 4 uint64)"""
        out = np.zeros(8, dtype=np.uint64)
        check(lib().gkrb200ec_g1_multiexp_device(self._h, slot, ctypes.c_void_p(d_scalars_ptr), n, form, _p(out)))
        return out

    def MultiExpPoints(self, points, scalars, form=SCALARS_REGULAR):
        """one-shot: exactly G1Affine.MultiExp(points, scalars, cfg)"""
        points, scalars = g1_array(points).reshape(-1, 8), fr_array(scalars).reshape(-1, 4)
        if points.shape[0] != scalars.shape[0]:
            raise ValueError("len(points) != len(scalars)")  # gnark-crypto returns an error here
        out = np.zeros(8, dtype=np.uint64)
        n = points.shape[0]
        check(lib().gkrb200ec_g1_multiexp_points(self._h, _p(points) if n else None, _p(scalars) if n else None, n, form, _p(out)))
        return out

    def Add(self, a, b):
        a, b = g1_array(a).reshape(8), g1_array(b).reshape(8)
        out = np.zeros(8, dtype=np.uint64)
        check(lib().gkrb200ec_g1_add(self._h, _p(a), _p(b), _p(out)))
        return out

    def InitialRandomnessHint(self, slot_pub, scalars_pub, slot_priv, scalars_priv, form=SCALARS_REGULAR):
        """prover/gadget/hints.go:162-192 -> (KrsGkrPriv (8,), initialRandomness (4,) regular form)"""
        sp, sq = fr_array(scalars_pub).reshape(-1, 4), fr_array(scalars_priv).reshape(-1, 4)
        krs_priv = np.zeros(8, dtype=np.uint64)
        rnd = np.zeros(4, dtype=np.uint64)
        check(lib().gkrb200ec_initial_randomness(self._h, slot_pub, _p(sp) if sp.shape[0] else None, sp.shape[0], slot_priv,
                                                 _p(sq) if sq.shape[0] else None, sq.shape[0], form, _p(krs_priv), _p(rnd)))
        return krs_priv, rnd

    # ---- G2 (prove.go:277: Bs.MultiExp(pk.G2.B, wireValuesB, ...))
    def SetBasesG2(self, slot, points):
        points = g2_array(points).reshape(-1, 16)
        check(lib().gkrb200ec_g2_set_bases(self._h, slot, _p(points) if points.shape[0] else None, points.shape[0]))

    def MultiExpG2(self, slot, scalars, form=SCALARS_REGULAR):
        """res.MultiExp(basesG2[slot][:len(scalars)], scalars, ecc.MultiExpConfig{}) -> (16,) G2Affine"""
        scalars = fr_array(scalars).reshape(-1, 4)
        out = np.zeros(16, dtype=np.uint64)
        check(lib().gkrb200ec_g2_multiexp(self._h, slot, _p(scalars) if scalars.shape[0] else None, scalars.shape[0], form, _p(out)))
        return out

    def MultiExpG2Device(self, slot, d_scalars_ptr, n, form=SCALARS_REGULAR):
        out = np.zeros(16, dtype=np.uint64)
        check(lib().gkrb200ec_g2_multiexp_device(self._h, slot, ctypes.c_void_p(d_scalars_ptr), n, form, _p(out)))
        return out

    def MultiExpPointsG2(self, points, scalars, form=SCALARS_REGULAR):
        points, scalars = g2_array(points).reshape(-1, 16), fr_array(scalars).reshape(-1, 4)
        if points.shape[0] != scalars.shape[0]:
            raise ValueError("len(points) != len(scalars)")
        out = np.zeros(16, dtype=np.uint64)
        n = points.shape[0]
        check(lib().gkrb200ec_g2_multiexp_points(self._h, _p(points) if n else None, _p(scalars) if n else None, n, form, _p(out)))
        return out

    def AddG2(self, a, b):
        a, b = g2_array(a).reshape(16), g2_array(b).reshape(16)
        out = np.zeros(16, dtype=np.uint64)
        check(lib().gkrb200ec_g2_add(self._h, _p(a), _p(b), _p(out)))
        return out

    # ---- fft.Domain / computeH
    def NewDomain(self, m):
        """fft.NewDomain(m, 1, true): the context's resident domain; returns its cardinality"""
        check(lib().gkrb200ec_fft_domain_init(self._h, int(m)))
        return int(lib().gkrb200ec_fft_domain_cardinality(self._h))

    @property
    def Cardinality(self):
        return int(lib().gkrb200ec_fft_domain_cardinality(self._h))

    def FFT(self, a, decimation, coset=0):
        """domain.FFT(a, decimation, coset) -> a new (n, 4) array (Go transforms in place)"""
        a = fr_array(a).reshape(-1, 4).copy()
        check(lib().gkrb200ec_fft(self._h, _p(a), a.shape[0], decimation, coset))
        return a

    def FFTInverse(self, a, decimation, coset=0):
        a = fr_array(a).reshape(-1, 4).copy()
        check(lib().gkrb200ec_fft_inverse(self._h, _p(a), a.shape[0], decimation, coset))
        return a

    def ComputeH(self, a, b, c):
        """computeH(a, b, c, &pk.Domain) -> (cardinality, 4) REGULAR-form words, coefficients in bit-reversed order"""
        a, b, c = (fr_array(v).reshape(-1, 4) for v in (a, b, c))
        if not (a.shape == b.shape == c.shape):
            raise ValueError("a, b, c must have the same length")
        h = np.zeros((self.Cardinality, 4), dtype=np.uint64)
        n_in = a.shape[0]
        check(lib().gkrb200ec_compute_h(self._h, _p(a) if n_in else None, _p(b) if n_in else None, _p(c) if n_in else None, n_in, _p(h), None))
        return h

    def ComputeHDevice(self, a, b, c):
        """the same, h left on the device: returns its raw device pointer (valid until the next FFT / ComputeH on this context),
        ready for MultiExpDevice(slot_of_pk_G1_Z, ptr, cardinality, SCALARS_REGULAR) -- krs2 of prove.go:221"""
        a, b, c = (fr_array(v).reshape(-1, 4) for v in (a, b, c))
        if not (a.shape == b.shape == c.shape):
            raise ValueError("a, b, c must have the same length")
        ptr = ctypes.c_void_p()
        n_in = a.shape[0]
        check(lib().gkrb200ec_compute_h(self._h, _p(a) if n_in else None, _p(b) if n_in else None, _p(c) if n_in else None, n_in, None,
                                        ctypes.byref(ptr)))
        return ptr.value

    def ComputeGroth16Proof(self, slots, points, a, b, c, wire_values_a, wire_values_b, r, s, form=SCALARS_REGULAR):
        """ComputeGroth16Proof(r1cs, pk, a, b, c, wireValues) (prover/gadget/prove.go:100-306) with the caller's r, s.
        slots = (slot of pk.G1.A, pk.G1.B, pk.G1.Z, pk.G2.B); points = dict g1_alpha, g1_beta, g1_delta (8,), g2_beta, g2_delta (16,);
        r, s: (4,) Montgomery fr.Element.  The context's fft domain must cover the constraints.  -> (Ar (8,), Bs (16,), Krs (8,))"""
        a, b, c = (fr_array(v).reshape(-1, 4) for v in (a, b, c))
        wa, wb = fr_array(wire_values_a).reshape(-1, 4), fr_array(wire_values_b).reshape(-1, 4)
        if not (a.shape == b.shape == c.shape):
            raise ValueError("a, b, c must have the same length")
        keep = {k: g1_array(points[k]).reshape(8) for k in ("g1_alpha", "g1_beta", "g1_delta")}
        keep.update({k: g2_array(points[k]).reshape(16) for k in ("g2_beta", "g2_delta")})
        pk = Groth16Pk(slots[0], slots[1], slots[2], slots[3], *[keep[k].ctypes.data for k in ("g1_alpha", "g1_beta", "g1_delta", "g2_beta", "g2_delta")])
        r, s = fr_array(r).reshape(4), fr_array(s).reshape(4)
        ar, bs, krs = np.zeros(8, dtype=np.uint64), np.zeros(16, dtype=np.uint64), np.zeros(8, dtype=np.uint64)
        n = a.shape[0]
        check(lib().gkrb200ec_groth16_prove(self._h, ctypes.byref(pk), _p(a) if n else None, _p(b) if n else None, _p(c) if n else None, n,
                                            _p(wa) if wa.shape[0] else None, wa.shape[0], _p(wb) if wb.shape[0] else None, wb.shape[0], form,
                                            _p(r), _p(s), _p(ar), _p(bs), _p(krs)))
        return ar, bs, krs

    def MultiExpSharded(self, slot, scalars_local, form=SCALARS_REGULAR, group=None):
        """One multi-exponentiation over several GPUs, one process per GPU (torch.distributed): rank r holds ITS slice of the bases in
        `slot` and passes the matching slice of the scalars; no data-path collective -- the only exchange is an all-gather of the
        partial sums (64 bytes per rank), which every rank then adds in rank order (a sum of points has one affine form: identical
        bytes on every rank, equal to the single-GPU result)."""
        import torch
        import torch.distributed as dist
        part = self.MultiExp(slot, scalars_local, form)
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return part
        t = torch.from_numpy(part.view(np.int64).copy())
        if dist.get_backend(group) == "nccl":
            t = t.to("cuda:%d" % self.device)
        parts = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, t, group=group)
        acc = np.zeros(8, dtype=np.uint64)
        for g in parts:
            acc = self.Add(acc, g.cpu().numpy().view(np.uint64))
        return acc

    def set_plan(self, window_bits=0, task_size=0):
        check(lib().gkrb200ec_set_plan(self._h, window_bits, task_size))

    def stats(self):
        st = EcStats()
        check(lib().gkrb200ec_get_stats(self._h, ctypes.byref(st)))
        return st
