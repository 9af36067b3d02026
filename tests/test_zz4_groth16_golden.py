"""tests/golden/groth16_side.json (written by oracle/gen_groth16_golden.py from the Python big-integer restatements) against
  * the C oracles (oracle/msm_oracle.c, oracle/fft_oracle.c),
  * the product's own host driver + kernel bodies on the CPU (tests/emu/ec_hostbuild.cpp: csrc/ec/ec.cu against a stand-in CUDA runtime),
  * and, with -m gpu, libgkrb200ec.so on the device.
The same file is what a maintainer with a Go toolchain feeds to gnark-crypto."""
import ctypes
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = json.load(open(os.path.join(ROOT, "tests", "golden", "groth16_side.json")))


def _mods():
    import cfft
    import cmsm
    cmsm.build()
    cfft.build()
    return cmsm, cfft


def _g1(cmsm, xy):
    return cmsm.point_from_ints((int(xy[0]), int(xy[1])))


def _g2(cmsm, p):
    vals = [int(p[0][0]), int(p[0][1]), int(p[1][0]), int(p[1][1])]
    return np.array(sum((cmsm.limbs(v * cmsm.RP % cmsm.P) for v in vals), []), dtype=np.uint64)


def _fr_reg(cmsm, vals):
    return cmsm.scalars_regular([int(v) for v in vals])


def _fr_mont(cmsm, vals):
    return cmsm.scalars_mont([int(v) for v in vals])


def _check_device_like(ectx, ec):
    """every golden vector through an EcContext (the host build on the CPU, the real library on the GPU)"""
    cmsm, cfft = _mods()
    g = G["g1_multiexp"]
    pts = np.array([_g1(cmsm, p) for p in g["points"]])
    want = _g1(cmsm, g["result"])
    assert np.array_equal(ectx.MultiExpPoints(pts, _fr_reg(cmsm, g["scalars"])), want)
    assert np.array_equal(ectx.MultiExpPoints(pts, _fr_mont(cmsm, g["scalars"]), ec.SCALARS_MONTGOMERY), want)
    for d in G["derive_randomness"]:
        pt = _g1(cmsm, d["point"])
        assert ec.RawBytes(pt).hex() == d["raw_bytes_hex"] and ec.LegacyKeccak256(bytes.fromhex(d["raw_bytes_hex"])).hex() == d["keccak256_hex"]
        assert cmsm.unlimbs(ec.DeriveRandomnessFromPoint(pt)) == int(d["randomness"])
    h = G["initial_randomness_hint"]
    ectx.SetBases(0, pts[:8])
    ectx.SetBases(1, pts[8:])
    sc = _fr_reg(cmsm, g["scalars"])
    krs_priv, rnd = ectx.InitialRandomnessHint(0, sc[:8], 1, sc[8:])
    assert np.array_equal(krs_priv, _g1(cmsm, h["krs_gkr_priv"])) and cmsm.unlimbs(rnd) == int(h["initial_randomness"])
    g2 = G["g2_multiexp"]
    pts2 = np.array([_g2(cmsm, p) for p in g2["points"]])
    assert np.array_equal(ectx.MultiExpPointsG2(pts2, _fr_reg(cmsm, g2["scalars"])), _g2(cmsm, g2["result"]))
    assert ectx.NewDomain(11) == int(G["domain_11"]["cardinality"]) == 16
    f = G["fft_16"]
    v = _fr_mont(cmsm, f["input"])
    assert np.array_equal(ectx.FFT(v, ec.DIF, 0), _fr_mont(cmsm, f["fft_dif_coset0"]))
    assert np.array_equal(ectx.FFT(v, ec.DIT, 1), _fr_mont(cmsm, f["fft_dit_coset1"]))
    assert np.array_equal(ectx.FFTInverse(v, ec.DIF, 1), _fr_mont(cmsm, f["fftinverse_dif_coset1"]))
    assert np.array_equal(ectx.FFTInverse(v, ec.DIT, 0), _fr_mont(cmsm, f["fftinverse_dit_coset0"]))
    c = G["compute_h_11"]
    a, b = _fr_mont(cmsm, c["a"]), _fr_mont(cmsm, c["b"])
    assert np.array_equal(ectx.ComputeH(a, b, _fr_mont(cmsm, c["c"])), _fr_reg(cmsm, c["h"]))
    assert np.array_equal(ectx.ComputeH(a, b, _fr_mont(cmsm, c["c_satisfied"])), _fr_reg(cmsm, c["h_satisfied"]))


def test_golden_file_is_self_consistent_and_matches_the_c_oracles():
    cmsm, cfft = _mods()
    assert G["compute_h_11"]["h_satisfied"] == G["compute_h_11"]["h_satisfied_by_long_division"]
    g = G["g1_multiexp"]
    pts = np.array([_g1(cmsm, p) for p in g["points"]])
    assert all(cmsm.is_on_curve(p) for p in pts)
    want = _g1(cmsm, g["result"])
    assert np.array_equal(cmsm.multiexp(pts, _fr_reg(cmsm, g["scalars"])), want)
    assert np.array_equal(cmsm.multiexp_buckets(pts, _fr_reg(cmsm, g["scalars"])), want)
    for d in G["derive_randomness"]:
        pt = _g1(cmsm, d["point"])
        assert cmsm.raw_bytes(pt).hex() == d["raw_bytes_hex"] and cmsm.keccak256(bytes.fromhex(d["raw_bytes_hex"])).hex() == d["keccak256_hex"]
        assert cmsm.unlimbs(cmsm.derive_randomness_from_point(pt)) == int(d["randomness"])
    g2 = G["g2_multiexp"]
    pts2 = np.array([_g2(cmsm, p) for p in g2["points"]])
    assert all(cmsm.g2_is_on_curve(p) for p in pts2)
    assert np.array_equal(cmsm.g2_multiexp(pts2, _fr_reg(cmsm, g2["scalars"])), _g2(cmsm, g2["result"]))
    gen, fine, ninv = cfft.domain(16)
    d = G["domain_11"]
    assert np.array_equal(gen, _fr_mont(cmsm, [d["generator"]])[0]) and np.array_equal(fine, _fr_mont(cmsm, [d["finer_generator"]])[0])
    assert np.array_equal(ninv, _fr_mont(cmsm, [d["cardinality_inv"]])[0])
    f = G["fft_16"]
    v = _fr_mont(cmsm, f["input"])
    assert np.array_equal(cfft.fft(v, cfft.DIF, 0), _fr_mont(cmsm, f["fft_dif_coset0"]))
    assert np.array_equal(cfft.fft(v, cfft.DIT, 1), _fr_mont(cmsm, f["fft_dit_coset1"]))
    assert np.array_equal(cfft.fft_inverse(v, cfft.DIF, 1), _fr_mont(cmsm, f["fftinverse_dif_coset1"]))
    assert np.array_equal(cfft.fft_inverse(v, cfft.DIT, 0), _fr_mont(cmsm, f["fftinverse_dit_coset0"]))
    c = G["compute_h_11"]
    assert np.array_equal(cfft.compute_h(_fr_mont(cmsm, c["a"]), _fr_mont(cmsm, c["b"]), _fr_mont(cmsm, c["c"]), 16), _fr_reg(cmsm, c["h"]))


def test_golden_vectors_through_the_host_build_of_the_driver():
    from gkrb200 import ec
    from test_ec_driver_cpu import build_hostbuild
    saved = ec._lib
    ec._lib = ec._bind(ctypes.CDLL(build_hostbuild()))
    try:
        with ec.EcContext(device=0) as ectx:
            _check_device_like(ectx, ec)
    finally:
        ec._lib = saved


@pytest.mark.gpu
def test_golden_vectors_on_the_device():
    from gkrb200 import ec
    with ec.EcContext(device=0) as ectx:
        _check_device_like(ectx, ec)
