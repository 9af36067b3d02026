"""GPU parity tests of the G1 multi-exponentiation path: libgkrb200ec.so on cuda:0, through its C ABI,
against the oracle (oracle/msm_oracle.c: one double-and-add per point, Jacobian coordinates -- nothing in common with the product's
bucket method in XYZZ coordinates).  Bit-exact: a sum of points has ONE affine form.

Mirrors the reference's call sites (it holds no unit test of MultiExp itself; its prover tests draw random keys):
    prover/gadget/hints.go:162-192   InitialRandomnessHint.Call  -> test_initial_randomness_hint_matches_oracle
    prover/gadget/hints.go:147-159   DeriveRandomnessFromPoint   -> same test (+ tests/test_msm_cpu.py on the host)
    prover/gadget/prove.go:76,91,... G1Affine.MultiExp           -> test_multiexp_matches_oracle, test_multiexp_full_size_closed_form
(The file name sorts last on purpose: this path was added after the GKR prover's tests and runs after them.)
"""
import ctypes
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cmsm():
    import cmsm as m
    m.build()
    return m


@pytest.fixture(scope="module")
def ecx():
    from gkrb200 import ec
    c = ec.EcContext(device=0)
    yield c
    c.close()


def _scalar_sets(cmsm, n, rng):
    q = cmsm.Q
    yield "random", [rng.randrange(q) for _ in range(n)]
    yield "small", [rng.randrange(1 << 16) for _ in range(n)]
    yield "edge", [[0, 1, q - 1, q - 2, 2, 1 << 253, (1 << 253) - 1, ((1 << 254) - 1) % q, q >> 1][i % 9] for i in range(n)]
    yield "all equal", [0x1F3C5A7799BBDDFF0123456789ABCDEF0FEDCBA987654321 % q] * n
    yield "all zero", [0] * n


def test_g1_add_on_the_device(cmsm, ecx):
    """G1Affine.Add (hints.go:184) incl. doubling, opposite points and infinity"""
    pts = cmsm.gen_points(6)
    zero = np.zeros(8, dtype=np.uint64)
    g = cmsm.generator()
    for a, b in [(pts[0], pts[1]), (pts[2], pts[2]), (pts[3], cmsm.neg(pts[3])), (zero, pts[4]), (pts[4], zero), (zero, zero), (g, g), (g, cmsm.neg(g))]:
        assert np.array_equal(ecx.Add(a, b), cmsm.add(a, b)), (cmsm.point_to_ints(a), cmsm.point_to_ints(b))


@pytest.mark.parametrize("n", [0, 1, 2, 7, 64, 300, 4096])
def test_multiexp_matches_oracle(cmsm, ecx, n):
    """G1Affine.MultiExp == sum_i s_i P_i for every scalar shape, forced window widths / task sizes (the result never depends on
    the plan), regular and Montgomery scalars, resident bases and the one-shot form"""
    from gkrb200 import ec
    rng = random.Random(100 + n)
    pts = cmsm.gen_points(n, a=rng.randrange(cmsm.Q), b=rng.randrange(cmsm.Q))
    if n >= 7:  # arbitrary bases: infinity, a repeated point, a pair of opposite points
        pts[1] = 0
        pts[3] = pts[2]
        pts[5] = cmsm.neg(pts[4])
    ecx.SetBases(0, pts)
    try:
        for name, vals in _scalar_sets(cmsm, n, rng):
            reg, mont = cmsm.scalars_regular(vals), cmsm.scalars_mont(vals)
            want = cmsm.multiexp(pts, reg) if n else np.zeros(8, dtype=np.uint64)
            for c, T in [(0, 0), (2, 0), (3, 1), (7, 2), (11, 0), (16, 0)]:
                ecx.set_plan(c, T)
                assert np.array_equal(ecx.MultiExp(0, reg), want), (name, c, T)
            ecx.set_plan(0, 0)
            assert np.array_equal(ecx.MultiExp(0, mont, ec.SCALARS_MONTGOMERY), want), name
            assert np.array_equal(ecx.MultiExpPoints(pts, reg), want), name
            if n >= 2:  # fewer scalars than bases: the prefix, as Go's slice pk.X[:len(scalars)]
                assert np.array_equal(ecx.MultiExp(0, reg[: n // 2]), cmsm.multiexp(pts[: n // 2], reg[: n // 2])), name
            if n:
                assert cmsm.is_on_curve(want)
    finally:
        ecx.set_plan(0, 0)
    if n:
        st = ecx.stats()
        assert st.launches_total > 0 and st.msm_calls > 0 and st.last_device_ms > 0
        assert st.last_n == (n // 2 if n >= 2 else n)


def test_multiexp_2pow16_and_skew(cmsm, ecx):
    """2^16 points: random scalars, and all scalars equal (one bucket per window holds every point and is cut into tasks) -- the
    shape of the reference's benchmark input, which hashes one value 2^k times"""
    n = 1 << 16
    rng = random.Random(16)
    pts = cmsm.gen_points(n)
    ecx.SetBases(1, pts)
    sc = cmsm.scalars_regular([rng.randrange(cmsm.Q) for _ in range(n)])
    want = cmsm.multiexp(pts, sc)
    got = ecx.MultiExp(1, sc)
    assert np.array_equal(got, want)
    assert np.array_equal(ecx.MultiExp(1, sc), got)  # entry order inside the buckets differs from run to run; the bytes do not
    same = cmsm.scalars_regular([0x2545F4914F6CDD1D2545F4914F6CDD1D2545F4914F6CDD1D % cmsm.Q] * n)
    assert np.array_equal(ecx.MultiExp(1, same), cmsm.multiexp(pts, same))
    ecx.SetBases(1, np.zeros((0, 8), dtype=np.uint64))


def test_initial_randomness_hint_matches_oracle(cmsm, ecx):
    """InitialRandomnessHint.Call (hints.go:162-192): KrsGkr = MultiExp(pubKGkr, pub) + MultiExp(privKGkrSigma, priv),
    initialRandomness = fr.SetBytes(Keccak256(KrsGkr.RawBytes())), KrsGkrPriv kept for the proof"""
    from gkrb200 import ec
    rng = random.Random(77)
    for n_pub, n_priv in [(1, 1), (3, 96), (192, 3 * 64), (3 * 1024, 5)]:
        pub = cmsm.gen_points(n_pub, a=rng.randrange(cmsm.Q), b=rng.randrange(cmsm.Q))
        priv = cmsm.gen_points(n_priv, a=rng.randrange(cmsm.Q), b=rng.randrange(cmsm.Q))
        sp = [rng.randrange(cmsm.Q) for _ in range(n_pub)]
        sq = [rng.randrange(cmsm.Q) for _ in range(n_priv)]
        ecx.SetBases(2, pub)
        ecx.SetBases(3, priv)
        want_priv, want_rnd = cmsm.initial_randomness(pub, cmsm.scalars_regular(sp), priv, cmsm.scalars_regular(sq))
        got_priv, got_rnd = ecx.InitialRandomnessHint(2, cmsm.scalars_regular(sp), 3, cmsm.scalars_regular(sq))
        assert np.array_equal(got_priv, want_priv) and np.array_equal(got_rnd, want_rnd)
        # the Go caller holds Montgomery fr.Elements and converts (hints.go:171); the device can do that conversion
        got_priv, got_rnd = ecx.InitialRandomnessHint(2, cmsm.scalars_mont(sp), 3, cmsm.scalars_mont(sq), ec.SCALARS_MONTGOMERY)
        assert np.array_equal(got_priv, want_priv) and np.array_equal(got_rnd, want_rnd)
        # pieces: DeriveRandomnessFromPoint of the sum
        krs = ecx.Add(ecx.MultiExp(2, cmsm.scalars_regular(sp)), got_priv)
        assert np.array_equal(ec.DeriveRandomnessFromPoint(krs), want_rnd)
    # opposite halves: KrsGkr = infinity, whose RawBytes are 0x40 00 .. 00
    pub = cmsm.gen_points(4)
    ecx.SetBases(2, pub)
    ecx.SetBases(3, pub)
    s = cmsm.scalars_regular([5, 6, 7, 8])
    s_neg = cmsm.scalars_regular([cmsm.Q - 5, cmsm.Q - 6, cmsm.Q - 7, cmsm.Q - 8])
    got_priv, got_rnd = ecx.InitialRandomnessHint(2, s, 3, s_neg)
    want_priv, want_rnd = cmsm.initial_randomness(pub, s, pub, s_neg)
    assert np.array_equal(got_priv, want_priv) and np.array_equal(got_rnd, want_rnd)
    assert np.array_equal(got_rnd, cmsm.derive_randomness_from_point(np.zeros(8, dtype=np.uint64)))


def test_multiexp_scalars_already_on_the_device(cmsm, ecx):
    """the GKR prover leaves its inputs/outputs in device memory as Montgomery fr.Elements: the multi-exponentiation reads them there"""
    import torch
    from gkrb200 import ec
    n = 1000
    rng = random.Random(5)
    pts = cmsm.gen_points(n)
    vals = [rng.randrange(cmsm.Q) for _ in range(n)]
    ecx.SetBases(4, pts)
    mont = cmsm.scalars_mont(vals)
    d = torch.from_numpy(mont.view(np.int64)).to("cuda:0")
    torch.cuda.synchronize()
    got = ecx.MultiExpDevice(4, d.data_ptr(), n, ec.SCALARS_MONTGOMERY)
    assert np.array_equal(got, cmsm.multiexp(pts, cmsm.scalars_regular(vals)))
    assert np.array_equal(d.cpu().numpy().view(np.uint64), mont)  # inputs untouched


def test_multiexp_argument_errors(cmsm, ecx):
    from gkrb200 import ec
    pts = cmsm.gen_points(3)
    ecx.SetBases(5, pts)
    sc = cmsm.scalars_regular([1, 2, 3])
    with pytest.raises(ec.GkrB200EcError) as e:
        ecx.MultiExp(5, cmsm.scalars_regular([1, 2, 3, 4]))  # more scalars than bases
    assert e.value.code == -1
    with pytest.raises(ec.GkrB200EcError):
        ecx.MultiExp(99, sc)
    with pytest.raises(ec.GkrB200EcError):
        ecx.MultiExp(6, sc)  # empty slot
    with pytest.raises(ec.GkrB200EcError):
        ecx.MultiExp(5, sc, 7)  # unknown scalar form
    bad = sc.copy()
    bad[1] = np.array(cmsm.limbs(cmsm.Q), dtype=np.uint64)  # q itself in regular form: fr.Element never holds it
    with pytest.raises(ec.GkrB200EcError) as e:
        ecx.MultiExp(5, bad)
    assert e.value.code == -1 and "not reduced" in str(e.value)
    assert np.array_equal(ecx.MultiExp(5, sc), cmsm.multiexp(pts, sc))  # the context survives an error
    with pytest.raises(ec.GkrB200EcError):
        ecx.set_plan(17, 0)
    with pytest.raises(ValueError):
        ecx.MultiExpPoints(pts, sc[:2])


def _limb_sums(arr):
    """(n, 4) uint64 regular-form scalars -> (sum_i s_i, sum_i i * s_i) as Python integers, through 16-bit pieces so that no
    64-bit partial sum overflows (i < 2^22, piece < 2^16, n <= 2^22 terms)"""
    idx = np.arange(arr.shape[0], dtype=np.uint64)
    tot = wtot = 0
    for j in range(4):
        for k in range(4):
            piece = (arr[:, j] >> np.uint64(16 * k)) & np.uint64(0xFFFF)
            tot += int(piece.sum()) << (64 * j + 16 * k)
            wtot += int((piece * idx).sum()) << (64 * j + 16 * k)
    return tot, wtot


def _add_limbs(x, y):
    """elementwise x + y of (n, 4) uint64 multi-limb integers whose sums stay below 2^256"""
    out = np.zeros_like(x)
    carry = np.zeros(x.shape[0], dtype=np.uint64)
    m32 = np.uint64(0xFFFFFFFF)
    for j in range(4):
        lo = (x[:, j] & m32) + (y[:, j] & m32) + carry
        hi = (x[:, j] >> np.uint64(32)) + (y[:, j] >> np.uint64(32)) + (lo >> np.uint64(32))
        out[:, j] = (lo & m32) | ((hi & m32) << np.uint64(32))
        carry = hi >> np.uint64(32)
    assert not carry.any()
    return out


@pytest.mark.parametrize("lg", [20, 22])
def test_multiexp_full_size_closed_form(cmsm, ecx, lg):
    """At the sizes of BASELINE.json's batches the oracle's point-by-point sum takes minutes; its bucket-method restatement takes
    seconds and is compared directly.  Independently of both, on the structured bases
    P_i = (a + i b) G the multi-exponentiation has a closed form: sum_i s_i P_i = (a sum s_i + b sum i s_i) G -- ONE scalar
    multiplication of the oracle.  Also linearity in the scalars on the same bases: MSM(t) + MSM(u) = MSM(t + u)."""
    n = 1 << lg
    a, b = 0x1234567, 0x9E3779B97F4A7C15
    q = cmsm.Q
    pts = cmsm.gen_points(n, a=a, b=b)
    ecx.SetBases(7, pts)
    rng = np.random.default_rng(lg)
    g = cmsm.generator()

    def rand252():  # uniform below 2^252 < q: canonical regular-form scalars without a reduction step
        w = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
        w[:, 3] >>= np.uint64(12)
        return w

    def closed(arr):
        tot, wtot = _limb_sums(arr)
        return cmsm.scalar_mul(g, (a * tot + b * wtot) % q)

    s = rand252()
    pyr = random.Random(lg)
    s[:2048] = cmsm.scalars_regular([pyr.randrange(q) for _ in range(2048)])  # the full range, incl. values above 2^253
    s[2048:2056] = cmsm.scalars_regular([0, 1, q - 1, q - 2, 1 << 253, (1 << 253) - 1, q >> 1, 2])
    ms = ecx.MultiExp(7, s)
    assert np.array_equal(ms, closed(s))
    assert cmsm.is_on_curve(ms)
    # and directly against the oracle's bucket-method restatement of gnark-crypto's MultiExp (seconds at this size; pinned to the
    # point-by-point oracle in tests/test_msm_cpu.py)
    assert np.array_equal(ms, cmsm.multiexp_buckets(pts, s))
    t, u = rand252(), rand252()
    mt, mu = ecx.MultiExp(7, t), ecx.MultiExp(7, u)
    assert np.array_equal(mt, closed(t)) and np.array_equal(mu, closed(u))
    assert np.array_equal(ecx.MultiExp(7, _add_limbs(t, u)), ecx.Add(mt, mu))
    stats = ecx.stats()
    assert stats.last_n == n and stats.last_c >= 10 and stats.last_device_ms > 0
    print("multiexp 2^%d: window %d bits, %d windows, %.2f ms on the device, workspace %.0f MiB"
          % (lg, stats.last_c, stats.last_windows, stats.last_device_ms, stats.workspace_bytes / 2**20))
    ecx.SetBases(7, np.zeros((0, 8), dtype=np.uint64))


# ------------------------------------------------------------------------------------------------ G2 (prove.go:277, Bs)
def test_g2_add_on_the_device(cmsm, ecx):
    pts = cmsm.g2_gen_points(6)
    zero = np.zeros(16, dtype=np.uint64)
    g = cmsm.g2_generator()
    for a, b in [(pts[0], pts[1]), (pts[2], pts[2]), (pts[3], cmsm.g2_neg(pts[3])), (zero, pts[4]), (pts[4], zero), (zero, zero), (g, g),
                 (g, cmsm.g2_neg(g))]:
        assert np.array_equal(ecx.AddG2(a, b), cmsm.g2_add(a, b))


@pytest.mark.parametrize("n", [0, 1, 2, 7, 64, 300, 4096])
def test_g2_multiexp_matches_oracle(cmsm, ecx, n):
    """G2Affine.MultiExp: the same kernels over Fp2 (16-word points), every scalar shape, forced plans, both scalar forms"""
    from gkrb200 import ec
    rng = random.Random(500 + n)
    pts = cmsm.g2_gen_points(n, a=rng.randrange(cmsm.Q), b=rng.randrange(cmsm.Q))
    if n >= 7:
        pts[1] = 0
        pts[3] = pts[2]
        pts[5] = cmsm.g2_neg(pts[4])
    ecx.SetBasesG2(9, pts)
    try:
        for name, vals in _scalar_sets(cmsm, n, rng):
            reg, mont = cmsm.scalars_regular(vals), cmsm.scalars_mont(vals)
            want = cmsm.g2_multiexp(pts, reg) if n else np.zeros(16, dtype=np.uint64)
            for c, T in [(0, 0), (2, 0), (3, 1), (7, 2), (16, 0)]:
                ecx.set_plan(c, T)
                assert np.array_equal(ecx.MultiExpG2(9, reg), want), (name, c, T)
            ecx.set_plan(0, 0)
            assert np.array_equal(ecx.MultiExpG2(9, mont, ec.SCALARS_MONTGOMERY), want), name
            assert np.array_equal(ecx.MultiExpPointsG2(pts, reg), want), name
            if n:
                assert cmsm.g2_is_on_curve(want)
    finally:
        ecx.set_plan(0, 0)
    if n:
        # a slot holds one kind of point
        with pytest.raises(ec.GkrB200EcError) as e:
            ecx.MultiExp(9, cmsm.scalars_regular([1]))
        assert e.value.code == -1 and "G2" in str(e.value)
    ecx.SetBasesG2(9, np.zeros((0, 16), dtype=np.uint64))


def test_g2_multiexp_2pow20_closed_form(cmsm, ecx):
    """2^20 points of G2 with a known discrete log: sum_i s_i P_i = (a sum s_i + b sum i s_i) G2gen, one scalar multiplication of the oracle"""
    n = 1 << 20
    a, b = 0x7654321, 0xD1B54A32D192ED03
    q = cmsm.Q
    pts = cmsm.g2_gen_points(n, a=a, b=b)
    ecx.SetBasesG2(10, pts)
    rng = np.random.default_rng(77)
    s = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    s[:, 3] >>= np.uint64(12)
    pyr = random.Random(20)
    s[:1024] = cmsm.scalars_regular([pyr.randrange(q) for _ in range(1024)])
    tot, wtot = _limb_sums(s)
    got = ecx.MultiExpG2(10, s)
    assert np.array_equal(got, cmsm.g2_scalar_mul(cmsm.g2_generator(), (a * tot + b * wtot) % q))
    assert cmsm.g2_is_on_curve(got)
    st = ecx.stats()
    print("G2 multiexp 2^20: window %d bits, %.2f ms on the device" % (st.last_c, st.last_device_ms))
    ecx.SetBasesG2(10, np.zeros((0, 16), dtype=np.uint64))
