"""CPU tests of the FFT half of the Groth16 prover (computeH, prover/gadget/prove.go:310-366).

  1. the ORACLE is pinned: the published 2^28-th root of unity of BN254 Fr (= 5^((q-1)/2^28), order exactly 2^28), the C restatement
     (oracle/fft_oracle.c: bit-reversal + textbook iterative transform) against the O(n^2) DEFINITIONS in oracle/pyref_fft.py, and
     the meaning of h: for a satisfied system it is the quotient (A B - C) / (X^n - 1), obtained there by long division, no FFT;
  2. the product's kernel bodies and drivers (gkr-mimc_b200/csrc/ec/ntt.cuh: in-register radix-8 / 4 / 2 passes, DIF and DIT,
     coset scalings, the fused computeH pipeline, the domain tables) compiled for the host by tests/emu/ec_emu.cpp and run launch
     by launch, forwards and backwards, against the oracle.
tests/test_zz1_ntt_gpu.py is the device parity test.
"""
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


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="module")
def cfft():
    import cfft as m
    m.build()
    return m


@pytest.fixture(scope="module")
def emu():
    deps = [EMU_SRC] + [os.path.join(EC_DIR, f) for f in ("field.cuh", "curve.cuh", "msm.cuh", "ntt.cuh")]
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(d) > os.path.getmtime(EMU_SO) for d in deps):
        os.makedirs(os.path.dirname(EMU_SO), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-Wextra", "-Werror", "-Wno-unknown-pragmas",
                               "-o", EMU_SO, EMU_SRC])
    L = ctypes.CDLL(EMU_SO)
    L.emu_ntt_rev.restype = ctypes.c_uint32
    return L


def _limbs(v):
    return [(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)]


def _unlimbs(row):
    return sum(int(x) << (64 * j) for j, x in enumerate(row))


def _to_mont(vals):
    import pyref_fft as pf
    r = (1 << 256) % pf.Q
    return np.array([_limbs(v * r % pf.Q) for v in vals], dtype=np.uint64).reshape(-1, 4)


def _from_mont(arr):
    import pyref_fft as pf
    rinv = pow((1 << 256) % pf.Q, -1, pf.Q)
    return [_unlimbs(row) * rinv % pf.Q for row in arr]


def _rand_elems(rng, n):
    """n canonical Montgomery images (any residue is one): uniform below 2^252 < q, plus the edge values"""
    x = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    x[:, 3] >>= np.uint64(12)
    import pyref_fft as pf
    edge = [0, 1, pf.Q - 1, pf.Q - 2, (1 << 256) % pf.Q, (1 << 253) + 7]
    for i, v in enumerate(edge[: max(0, min(len(edge), n - 1))]):
        x[(i * 7 + 1) % n] = np.array(_limbs(v), dtype=np.uint64)
    return x


# ------------------------------------------------------------------------------------------------ 1. the oracle is pinned
def test_root_of_unity_known_answers(cfft):
    import pyref_fft as pf
    g = pf.ROOT_OF_UNITY
    assert pow(g, 1 << 28, pf.Q) == 1 and pow(g, 1 << 27, pf.Q) == pf.Q - 1  # order exactly 2^28
    assert (pf.Q - 1) % (1 << 28) == 0 and (pf.Q - 1) % (1 << 29) != 0     # the 2-adicity of BN254 Fr
    assert pow(5, (pf.Q - 1) >> 28, pf.Q) == g                              # 5 generates Fr*; this is gnark-crypto's root
    for m in (1, 2, 5, 1 << 10, (1 << 20) - 3, 1 << 22):
        dom = pf.Domain(m)
        assert dom.n >= m and dom.n & (dom.n - 1) == 0 and dom.n < 2 * max(m, 1)
        assert pow(dom.generator, dom.n, pf.Q) == 1 and (dom.n == 1 or pow(dom.generator, dom.n // 2, pf.Q) == pf.Q - 1)
        assert pow(dom.finer_generator, 2, pf.Q) == dom.generator and pow(dom.finer_generator, dom.n, pf.Q) == pf.Q - 1
        gen, fine, ninv = cfft.domain(dom.n)
        assert _from_mont([gen, fine, ninv]) == [dom.generator, dom.finer_generator, dom.cardinality_inv]


def test_c_oracle_matches_the_definitions(cfft):
    import pyref_fft as pf
    rng = random.Random(2)
    for n in (1, 2, 4, 8, 32):
        dom = pf.Domain(n)
        v = [rng.randrange(pf.Q) for _ in range(n)]
        for dec in (pf.DIT, pf.DIF):
            for cs in (0, 1):
                assert _from_mont(cfft.fft(_to_mont(v), dec, cs)) == pf.fft(dom, v, dec, cs), (n, dec, cs)
                assert _from_mont(cfft.fft_inverse(_to_mont(v), dec, cs)) == pf.fft_inverse(dom, v, dec, cs), (n, dec, cs)
    for m in (1, 3, 8, 13, 20):
        dom = pf.Domain(m)
        a, b, c = ([rng.randrange(pf.Q) for _ in range(m)] for _ in range(3))
        assert [_unlimbs(r) for r in cfft.compute_h(_to_mont(a), _to_mont(b), _to_mont(c))] == pf.compute_h(a, b, c, dom), m


def test_compute_h_is_the_quotient_polynomial(cfft):
    """for a satisfied system (a[i] b[i] = c[i] for every constraint; the padding is 0 * 0 = 0) h is (A B - C) / (X^n - 1): pins
    the coset (u^n = -1, hence the -2), the scalings and the bit-reversed output order with no FFT on the checking side"""
    import pyref_fft as pf
    rng = random.Random(4)
    for m in (1, 2, 5, 8, 11, 16):
        dom = pf.Domain(m)
        a = [rng.randrange(pf.Q) for _ in range(m)]
        b = [rng.randrange(pf.Q) for _ in range(m)]
        c = [x * y % pf.Q for x, y in zip(a, b)]
        want = pf.compute_h_by_division(a, b, c, dom)
        assert pf.compute_h(a, b, c, dom) == want
        assert [_unlimbs(r) for r in cfft.compute_h(_to_mont(a), _to_mont(b), _to_mont(c))] == want
        if dom.n > 1:
            assert any(want)


# ------------------------------------------------------------------------------------------------ 2. the kernel bodies on the host
def test_emulated_domain_and_bit_reversal(cfft, emu):
    out = np.zeros(16, dtype=np.uint64)
    import pyref_fft as pf
    for log in range(0, 27):
        assert emu.emu_fft_domain(log, _p(out)) == 0
        gen, fine, ninv = cfft.domain(1 << log)
        assert np.array_equal(out[:4], gen) and np.array_equal(out[4:8], fine) and np.array_equal(out[8:12], ninv), log
        assert _from_mont([out[12:16]])[0] == pow(pf.Q - 2, -1, pf.Q)
    assert emu.emu_fft_domain(27, _p(out)) == -1
    for log in (1, 2, 5, 11, 12, 22, 26):
        for x in (0, 1, 2, (1 << log) - 1, 0x2545F491 & ((1 << log) - 1)):
            assert emu.emu_ntt_rev(x, log) == pf.rev(x, log), (x, log)


@pytest.mark.parametrize("log", list(range(0, 10)) + [12])
def test_emulated_fft_matches_oracle(cfft, emu, log):
    """domain.FFT / FFTInverse, DIF and DIT, coset 0 and 1: all 8 variants, both launch orders; 2^log covers every mix of radix-8, -4
    and -2 passes (log mod 3) and the boundary of the split coset tables (2^11)"""
    n = 1 << log
    rng = np.random.default_rng(100 + log)
    v = _rand_elems(rng, n)
    for dec in (0, 1):
        for cs in (0, 1):
            for inv in (0, 1):
                want = (cfft.fft_inverse if inv else cfft.fft)(v, dec, cs)
                for order in (0, 1):
                    a = v.copy()
                    assert emu.emu_fft(_p(a), log, dec, cs, inv, order) >= 0
                    assert np.array_equal(a, want), (log, dec, cs, inv, order)
    # round trips as computeH chains them: DIF then DIT, no bit-reversal pass in between
    a = v.copy()
    emu.emu_fft(_p(a), log, 1, 1, 0, 0)
    emu.emu_fft(_p(a), log, 0, 1, 1, 0)
    assert np.array_equal(a, v)


@pytest.mark.parametrize("log", [0, 1, 2, 3, 4, 7, 11, 13])
def test_emulated_compute_h_matches_oracle(cfft, emu, log):
    n = 1 << log
    rng = np.random.default_rng(200 + log)
    for m in sorted({n, max(1, n - 3), n // 2 + 1}):
        a, b, c = _rand_elems(rng, m), _rand_elems(rng, m), _rand_elems(rng, m)
        want = cfft.compute_h(a, b, c, n)
        for order in (0, 1):
            h = np.full((n, 4), 0xA5A5A5A5A5A5A5A5, dtype=np.uint64)
            assert emu.emu_compute_h(_p(a), _p(b), _p(c), ctypes.c_size_t(m), log, order, _p(h)) > 0
            assert np.array_equal(h, want), (log, m, order)
    assert emu.emu_compute_h(_p(a), _p(b), _p(c), ctypes.c_size_t(n + 1), log, 0, _p(h)) == -1


def test_emulated_compute_h_of_a_satisfied_system(cfft, emu):
    """the product pipeline against the quotient by long division (Python, no FFT anywhere on the checking side)"""
    import pyref_fft as pf
    rng = random.Random(8)
    for m in (5, 16, 27):
        dom = pf.Domain(m)
        a = [rng.randrange(pf.Q) for _ in range(m)]
        b = [rng.randrange(pf.Q) for _ in range(m)]
        c = [x * y % pf.Q for x, y in zip(a, b)]
        h = np.zeros((dom.n, 4), dtype=np.uint64)
        am, bm, cm = _to_mont(a), _to_mont(b), _to_mont(c)
        assert emu.emu_compute_h(_p(am), _p(bm), _p(cm), ctypes.c_size_t(m), dom.log, 0, _p(h)) > 0
        assert [_unlimbs(r) for r in h] == pf.compute_h_by_division(a, b, c, dom)


def test_ec_fft_entry_points_fail_loudly_without_a_domain_or_device():
    from gkrb200 import ec
    L = ec.lib()
    a = np.zeros((4, 4), dtype=np.uint64)
    assert L.gkrb200ec_fft(None, _p(a), 4, 0, 0) == -1
    assert L.gkrb200ec_fft_domain_init(None, 4) == -1
    assert L.gkrb200ec_compute_h(None, _p(a), _p(a), _p(a), 4, _p(a), None) == -1
    assert L.gkrb200ec_fft_domain_cardinality(None) == 0
