"""CPU test of the ComputeGroth16Proof sequencing (gkr-mimc_b200/csrc/ec/groth16.hpp; prover/gadget/prove.go:100-306): the product's
composition run over the emulated kernels (tests/emu/ec_emu.cpp) against the oracle's own composition of the oracle's own pieces
(oracle/cgroth16.py).  tests/test_zz3_groth16_gpu.py runs the same comparison through the C ABI on the device."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_SRC = os.path.join(ROOT, "tests", "emu", "ec_emu.cpp")
EMU_SO = os.path.join(ROOT, "tests", "emu", "_build", "libecemu.so")
EC_DIR = os.path.join(ROOT, "gkr-mimc_b200", "csrc", "ec")


class EmuGroth16In(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in ("g1_a", "g1_b", "g1_z", "g2_b", "alpha", "beta", "delta", "beta2", "delta2", "a", "b", "c", "wa", "wb")] + \
               [("n_constraints", ctypes.c_uint64), ("n_a", ctypes.c_uint64), ("n_b", ctypes.c_uint64), ("log_n", ctypes.c_uint32),
                ("scalars_mont", ctypes.c_int)]


@pytest.fixture(scope="module")
def emu():
    deps = [EMU_SRC] + [os.path.join(EC_DIR, f) for f in os.listdir(EC_DIR) if f.endswith((".cuh", ".hpp"))]
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(d) > os.path.getmtime(EMU_SO) for d in deps):
        os.makedirs(os.path.dirname(EMU_SO), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-Wextra", "-Werror", "-Wno-unknown-pragmas",
                               "-o", EMU_SO, EMU_SRC])
    L = ctypes.CDLL(EMU_SO)
    L.emu_groth16.restype = ctypes.c_int
    return L


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p).value


@pytest.mark.parametrize("m,n_a,n_b,mont", [(5, 7, 6, 0), (16, 20, 13, 1), (100, 90, 77, 0)])
def test_emulated_groth16_composition_matches_oracle(emu, m, n_a, n_b, mont):
    import cfft
    import cgroth16
    import cmsm
    cmsm.build()
    cfft.build()
    n = cfft.next_pow2(m)
    rng = random.Random(m)
    pk = cgroth16.synthetic_proving_key(n_a, n_b, n, seed=m)
    rand_fr = lambda k: cmsm.scalars_mont([rng.randrange(cmsm.Q) for _ in range(k)])
    a, b, c = rand_fr(m), rand_fr(m), rand_fr(m)
    wa_vals = [rng.randrange(cmsm.Q) for _ in range(n_a)]
    wb_vals = [rng.randrange(cmsm.Q) for _ in range(n_b)]
    r, s = rng.randrange(cmsm.Q), rng.randrange(cmsm.Q)
    want_ar, want_bs, want_krs = cgroth16.compute_groth16_proof(pk, a, b, c, cmsm.scalars_regular(wa_vals), cmsm.scalars_regular(wb_vals), r, s, n)
    wa = (cmsm.scalars_mont if mont else cmsm.scalars_regular)(wa_vals)
    wb = (cmsm.scalars_mont if mont else cmsm.scalars_regular)(wb_vals)
    keep = [np.ascontiguousarray(pk[k]) for k in ("g1_a", "g1_b", "g1_z", "g2_b", "g1_alpha", "g1_beta", "g1_delta", "g2_beta", "g2_delta")]
    inp = EmuGroth16In(*[_p(x) for x in keep], _p(a), _p(b), _p(c), _p(wa), _p(wb), m, n_a, n_b, n.bit_length() - 1, mont)
    ar, bs, krs = np.zeros(8, dtype=np.uint64), np.zeros(16, dtype=np.uint64), np.zeros(8, dtype=np.uint64)
    rm, sm = cgroth16.fr_mont(r), cgroth16.fr_mont(s)
    rc = emu.emu_groth16(ctypes.byref(inp), ctypes.c_void_p(_p(rm)), ctypes.c_void_p(_p(sm)), ctypes.c_void_p(_p(ar)), ctypes.c_void_p(_p(bs)),
                         ctypes.c_void_p(_p(krs)))
    assert rc == 0
    assert np.array_equal(ar, want_ar) and np.array_equal(bs, want_bs) and np.array_equal(krs, want_krs)
    assert cmsm.is_on_curve(ar) and cmsm.is_on_curve(krs) and cmsm.g2_is_on_curve(bs)
