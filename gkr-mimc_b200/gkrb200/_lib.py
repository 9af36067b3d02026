"""Loader for libgkrb200.so (the C-ABI CUDA library).  Fails loudly if the library is missing: there is
no Python/CPU fallback for any device operation."""
import ctypes
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_PKG)  # gkr-mimc_b200/
_VARIANT = os.environ.get("GKRB200_LIB_VARIANT", "")  # "" (product build) | "kara" | "mixed": A/B builds of the multiplier (make variants)
SO_PATH = os.path.join(ROOT, "libgkrb200%s.so" % ("_" + _VARIANT if _VARIANT else ""))


class Stats(ctypes.Structure):
    _fields_ = [
        ("launches_total", ctypes.c_uint64),
        ("launches", ctypes.c_uint64 * 8),
        ("kernel_ms", ctypes.c_double * 8),
        ("transcript_ms", ctypes.c_double),
        ("wait_ms", ctypes.c_double),
        ("comm_ms", ctypes.c_double),
        ("rounds", ctypes.c_uint64),
        ("fr_mul_assign", ctypes.c_uint64),
        ("fr_mul_round", ctypes.c_uint64),
        ("bytes_round", ctypes.c_uint64),
        ("h2d_bytes", ctypes.c_uint64),
        ("d2h_bytes", ctypes.c_uint64),
        ("follow_wait_ms", ctypes.c_double),
    ]


def build(verbose=False):
    """Compile the library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-C", ROOT] + ([] if verbose else ["-s"]))
    return SO_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            "gkrb200: %s is missing. Build it with `make -C %s` (or __graft_entry__.build()); "
            "there is no CPU fallback." % (SO_PATH, ROOT))
    L = ctypes.CDLL(SO_PATH)
    vp, sz, i32, u32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32
    L.gkrb200_last_error.restype = ctypes.c_char_p
    L.gkrb200_version.restype = ctypes.c_char_p
    L.gkrb200_init.argtypes = [ctypes.POINTER(vp), i32, i32, vp]
    L.gkrb200_init_shard.argtypes = [ctypes.POINTER(vp), i32, i32, vp, i32]
    L.gkrb200_free.argtypes = [vp]
    L.gkrb200_free.restype = None
    L.gkrb200_comm_unique_id.argtypes = [vp]
    L.gkrb200_comm_init.argtypes = [vp, i32, i32, vp]
    L.gkrb200_comm_set_leader.argtypes = [vp, i32]
    L.gkrb200_comm_exchange_mode.argtypes = [vp]
    L.gkrb200_mimc_assign.argtypes = [vp, vp, vp, sz, vp]
    L.gkrb200_mimc_assign_device.argtypes = [vp, vp, vp, sz]
    L.gkrb200_assign_layer_to_host.argtypes = [vp, i32, vp, sz]
    L.gkrb200_gkr_prove_mimc.argtypes = [vp, vp, i32, vp, u32]
    L.gkrb200_proof_vec_len.argtypes = [i32]
    L.gkrb200_proof_vec_len.restype = sz
    L.gkrb200_sumcheck_prove.argtypes = [vp, vp, vp, i32, vp, sz, vp, sz, i32, vp, vp, vp, vp]
    L.gkrb200_sumcheck_prove_device.argtypes = [vp, vp, vp, i32, vp, sz, vp, sz, i32, vp, vp, vp, vp]
    L.gkrb200_check_mimc_circuit.argtypes = [i32, vp, vp, vp, vp]
    L.gkrb200_eq_table.argtypes = [vp, vp, sz, i32, vp, vp]
    L.gkrb200_fold.argtypes = [vp, vp, sz, vp, vp]
    L.gkrb200_round_eval.argtypes = [vp, vp, vp, vp, sz, i32, vp, vp]
    L.gkrb200_fr_batch.argtypes = [vp, i32, vp, vp, sz, vp]
    L.gkrb200_mimc_hash.argtypes = [vp, sz, vp]
    L.gkrb200_interpolate.argtypes = [vp, sz, vp]
    L.gkrb200_to_montgomery.argtypes = [vp, sz, vp]
    L.gkrb200_from_montgomery.argtypes = [vp, sz, vp]
    L.gkrb200_mimc_ark.argtypes = [i32, vp]
    L.gkrb200_const_mul_table.argtypes = [vp, vp]
    L.gkrb200_eval_univariate.argtypes = [vp, sz, vp, vp]
    L.gkrb200_eval_eq.argtypes = [vp, vp, sz, vp]
    L.gkrb200_fr_scalar.argtypes = [i32, vp, vp, vp]
    L.gkrb200_sumcheck_verify.argtypes = [vp, sz, vp, i32, i32, vp, vp, vp]
    L.gkrb200_stats_reset.argtypes = [vp]
    L.gkrb200_stats_get.argtypes = [vp, ctypes.POINTER(Stats)]
    L.gkrb200_set_profiling.argtypes = [vp, i32]
    L.gkrb200_mimc_assign_ex.argtypes = [vp, vp, vp, sz, vp, u32]
    L.gkrb200_convert.argtypes = [vp, vp, sz, vp, i32]
    L.gkrb200_mle_evaluate.argtypes = [vp, vp, sz, vp, vp]
    L.gkrb200_assign_layer_evaluate.argtypes = [vp, i32, vp, i32, vp]
    L.gkrb200_gkr_verify_mimc.argtypes = [vp, vp, i32, vp, u32]
    L.gkrb200_gkr_verify_mimc_io.argtypes = [vp, vp, i32, vp, u32, vp, vp, vp]
    L.gkrb200_set_option.argtypes = [vp, i32, ctypes.c_long]
    L.gkrb200_microbench.argtypes = [vp, i32, i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    _lib = L
    return L


class GkrB200Error(RuntimeError):
    """Raised where the reference panics (sumcheck/prover.go:54,:114; gkr/prover.go:84; poly/pool.go:71)."""

    def __init__(self, code, msg):
        super().__init__("gkrb200 error %d: %s" % (code, msg))
        self.code = code


def check(rc):
    if rc != 0:
        raise GkrB200Error(rc, lib().gkrb200_last_error().decode("utf-8", "replace"))
