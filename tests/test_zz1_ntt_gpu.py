"""GPU parity tests of the FFT half of the Groth16 prover (computeH, prover/gadget/prove.go:310-366):
libgkrb200ec.so on cuda:0, through its C ABI, against the oracle (oracle/fft_oracle.c: explicit bit reversal + textbook iterative
transform; tests/test_ntt_cpu.py pins it to the O(n^2) definitions and to the quotient by long division).  Bit-exact.

    fft.Domain.FFT / FFTInverse (DIF, DIT, coset 0 / 1)  -> test_fft_variants_match_oracle
    computeH(a, b, c, &pk.Domain)                        -> test_compute_h_matches_oracle, test_compute_h_full_size
    krs2.MultiExp(pk.G1.Z, h, ...) (prove.go:221)        -> test_compute_h_result_feeds_the_multiexp_on_the_device
(The file name sorts last on purpose: this path was added after the GKR prover's tests and runs after them.)
"""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfft():
    import cfft as m
    m.build()
    return m


@pytest.fixture(scope="module")
def ecx():
    from gkrb200 import ec
    c = ec.EcContext(device=0)
    yield c
    c.close()


def _limbs(v):
    return [(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)]


def _rand_elems(rng, n):
    """n canonical fr.Element images: uniform below 2^252 < q, with the edge values mixed in"""
    import pyref_fft as pf
    x = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    x[:, 3] >>= np.uint64(12)
    edge = [0, 1, pf.Q - 1, pf.Q - 2, (1 << 256) % pf.Q, (1 << 253) + 7]
    for i, v in enumerate(edge[: max(0, min(len(edge), n - 1))]):
        x[(i * 7 + 1) % n] = np.array(_limbs(v), dtype=np.uint64)
    return x


@pytest.mark.parametrize("log", [0, 1, 2, 3, 4, 5, 6, 10, 11, 12, 16])
def test_fft_variants_match_oracle(cfft, ecx, log):
    """all 8 variants; 2^log covers every mix of radix-8 / 4 / 2 passes and the 2^11 boundary of the split coset tables"""
    from gkrb200 import ec
    n = 1 << log
    assert ecx.NewDomain(n) == n and ecx.Cardinality == n
    rng = np.random.default_rng(100 + log)
    v = _rand_elems(rng, n)
    for dec in (ec.DIT, ec.DIF):
        for cs in (0, 1):
            assert np.array_equal(ecx.FFT(v, dec, cs), cfft.fft(v, dec, cs)), (log, dec, cs)
            assert np.array_equal(ecx.FFTInverse(v, dec, cs), cfft.fft_inverse(v, dec, cs)), (log, dec, cs)
    # chained as computeH does: DIF forward then DIT inverse, no bit-reversal pass in between
    assert np.array_equal(ecx.FFTInverse(ecx.FFT(v, ec.DIF, 1), ec.DIT, 1), v)
    st = ecx.stats()
    assert st.fft_calls > 0 and st.launches_total > 0


@pytest.mark.parametrize("m", [1, 2, 5, 1000, 4096, (1 << 16) - 3, 1 << 16])
def test_compute_h_matches_oracle(cfft, ecx, m):
    """computeH on m constraints (zero padding to the domain's cardinality happens on the device), arbitrary a, b, c"""
    n = ecx.NewDomain(m)
    assert n == cfft.next_pow2(m)
    rng = np.random.default_rng(m)
    a, b, c = _rand_elems(rng, m), _rand_elems(rng, m), _rand_elems(rng, m)
    want = cfft.compute_h(a, b, c, n)
    got = ecx.ComputeH(a, b, c)
    assert got.shape == (n, 4) and np.array_equal(got, want)
    assert np.array_equal(ecx.ComputeH(a, b, c), got)
    # a satisfied system: h is the quotient (A B - C) / (X^n - 1)
    c = cfft.mul_elementwise(a, b)
    got = ecx.ComputeH(a, b, c)
    assert np.array_equal(got, cfft.compute_h(a, b, c, n))
    z = _rand_elems(rng, 8)[7]
    assert cfft.quotient_identity_holds(a, b, c, got, n, z)


@pytest.mark.parametrize("lg", [20, 22])
def test_compute_h_full_size(cfft, ecx, lg):
    """at BASELINE.json's batch sizes: every word against the oracle at 2^20; at 2^22 (where the oracle's seven transforms take a
    minute) the defining identity H(z) (z^n - 1) = A(z) B(z) - C(z) of a satisfied system at a random point -- O(n) oracle work,
    wrong with probability < 2^-230 if any coefficient of h is off -- plus exact equality on a perturbed copy being REJECTED"""
    n = 1 << lg
    m = n - 5
    assert ecx.NewDomain(m) == n
    rng = np.random.default_rng(lg)
    a, b = _rand_elems(rng, m), _rand_elems(rng, m)
    c = cfft.mul_elementwise(a, b)
    h = ecx.ComputeH(a, b, c)
    z = _rand_elems(rng, 8)[7]
    assert cfft.quotient_identity_holds(a, b, c, h, n, z)
    bad = h.copy()
    bad[n // 3, 1] ^= np.uint64(1 << 17)
    assert not cfft.quotient_identity_holds(a, b, c, bad, n, z)
    if lg <= 20:
        assert np.array_equal(h, cfft.compute_h(a, b, c, n))
    st = ecx.stats()
    print("computeH 2^%d: %.2f ms on the device (7 transforms + 5 elementwise passes)" % (lg, st.last_fft_device_ms))
    assert st.last_fft_device_ms > 0


def test_compute_h_result_feeds_the_multiexp_on_the_device(cfft, ecx):
    """krs2.MultiExp(pk.G1.Z, h, ...) (prove.go:221): h stays on the device between computeH and the multi-exponentiation"""
    import cmsm
    from gkrb200 import ec
    cmsm.build()
    m = 1000
    n = ecx.NewDomain(m)
    rng = np.random.default_rng(9)
    a, b, c = _rand_elems(rng, m), _rand_elems(rng, m), _rand_elems(rng, m)
    zpts = cmsm.gen_points(n, a=0xABCDEF, b=0x13579BDF)
    ecx.SetBases(8, zpts)
    h = ecx.ComputeH(a, b, c)
    want = cmsm.multiexp(zpts, h)
    assert np.array_equal(ecx.MultiExp(8, h), want)
    d_h = ecx.ComputeHDevice(a, b, c)
    assert d_h
    assert np.array_equal(ecx.MultiExpDevice(8, d_h, n, ec.SCALARS_REGULAR), want)
    ecx.SetBases(8, np.zeros((0, 8), dtype=np.uint64))


def test_fft_argument_errors(cfft):
    from gkrb200 import ec
    with ec.EcContext(device=0) as c:
        v = np.zeros((8, 4), dtype=np.uint64)
        with pytest.raises(ec.GkrB200EcError) as e:
            c.FFT(v, ec.DIF, 0)  # no domain yet
        assert e.value.code == -1 and "domain" in str(e.value)
        with pytest.raises(ec.GkrB200EcError):
            c.NewDomain(0)
        with pytest.raises(ec.GkrB200EcError):
            c.NewDomain((1 << 26) + 1)
        assert c.NewDomain(8) == 8
        with pytest.raises(ec.GkrB200EcError):
            c.FFT(np.zeros((4, 4), dtype=np.uint64), ec.DIF, 0)  # not the domain's cardinality
        with pytest.raises(ec.GkrB200EcError):
            c.FFT(v, 2, 0)
        with pytest.raises(ec.GkrB200EcError):
            c.FFT(v, ec.DIF, 2)  # depth-1 domain: cosets 0 and 1
        with pytest.raises(ec.GkrB200EcError):
            c.ComputeH(np.zeros((9, 4), dtype=np.uint64), np.zeros((9, 4), dtype=np.uint64), np.zeros((9, 4), dtype=np.uint64))
        with pytest.raises(ValueError):
            c.ComputeH(v, v, v[:4])
        assert not c.FFT(v, ec.DIF, 0).any()  # the context survives the errors; the transform of zero is zero
        assert c.NewDomain(3) == 4 and c.Cardinality == 4  # replacing the domain
