"""CPU tests of the G1 multi-exponentiation path (InitialRandomnessHint: prover/gadget/hints.go:147-192).

Three layers, none of which needs a GPU:
  1. the ORACLE is pinned: published Keccak-256 and BN254 known answers, and the C restatement (oracle/msm_oracle.c, bit-by-bit
     double-and-add in Jacobian coordinates) agrees with the independent Python big-integer one (oracle/pyref_msm.py);
  2. the product's kernel bodies (gkr-mimc_b200/csrc/ec/*.cuh: field arithmetic, XYZZ group law with every exceptional case, digit
     decomposition, counting sort, task splitting, bucket / chunk / window reduction, the launch sequence of msm_enqueue) are
     compiled for the host by tests/emu/ec_emu.cpp and compared with the oracle -- each "launch" run as a loop, forwards and
     backwards; on the host the carry chains are plain C++, on the device they are the inline-PTX primitives of fr_device.cuh;
  3. libgkrb200ec.so loads, exports every symbol include/gkrb200_ec.h declares, its host-only entry points (RawBytes, legacy
     Keccak-256, DeriveRandomnessFromPoint) agree with the oracle, and the device entry points fail loudly without a GPU.
tests/test_zz2_msm_gpu.py is the device parity test proper.
"""
import ctypes
import os
import random
import re
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
def cmsm():
    import cmsm as m
    m.build()
    return m


@pytest.fixture(scope="module")
def emu():
    deps = [EMU_SRC] + [os.path.join(EC_DIR, f) for f in ("field.cuh", "curve.cuh", "msm.cuh")]
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(d) > os.path.getmtime(EMU_SO) for d in deps):
        os.makedirs(os.path.dirname(EMU_SO), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-Wextra", "-Werror", "-Wno-unknown-pragmas",
                               "-o", EMU_SO, EMU_SRC])
    L = ctypes.CDLL(EMU_SO)
    L.emu_msm.restype = ctypes.c_int
    L.emu_g2_msm.restype = ctypes.c_int
    return L


def _emu_msm(emu, points, scalars, mont=0, c=0, T=0, reverse=0):
    points, scalars = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 8), np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros(16, dtype=np.uint64)
    plan = np.zeros(8, dtype=np.uint32)
    rc = emu.emu_msm(_p(points), _p(scalars), ctypes.c_size_t(points.shape[0]), mont, c, T, reverse, _p(out), _p(plan))
    return rc, out, plan


# ------------------------------------------------------------------------------------------------ 1. the oracle is pinned
KECCAK_KAT = [  # legacy Keccak-256 (the Ethereum hash): published known answers
    (b"", "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"),
    (b"abc", "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"),
    (b"hello world", "47173285a8d7341e5e972fc677286384f802f8ef42a5ec5f03bbfa254cb01fad"),
]
G2X = 1368015179489954701390400359078579693043519447331113978918064868415326638035  # 2 * (1, 2) on BN254 (EIP-196 vectors)
G2Y = 9918110051302171585080402603319702774565515993150576347155970296011118125764


def test_oracle_keccak_known_answers(cmsm):
    import pyref_msm as pr
    for data, want in KECCAK_KAT:
        assert pr.keccak256(data).hex() == want
        assert cmsm.keccak256(data).hex() == want
    assert pr.keccak256(b"transfer(address,uint256)")[:4].hex() == "a9059cbb"  # the ERC-20 selector
    rng = random.Random(5)
    for n in list(range(0, 10)) + [63, 64, 65, 134, 135, 136, 137, 271, 272, 273, 500]:  # around the 136-byte rate
        data = bytes(rng.randrange(256) for _ in range(n))
        assert cmsm.keccak256(data) == pr.keccak256(data), n


def test_oracle_bn254_known_answers(cmsm):
    import pyref_msm as pr
    assert pr.is_on_curve(pr.G1) and pr.add(pr.G1, pr.G1) == (G2X, G2Y)
    assert pr.mul(pr.Q, pr.G1) == pr.INF and pr.mul(pr.Q - 1, pr.G1) == pr.neg(pr.G1)
    g = cmsm.generator()
    assert cmsm.point_to_ints(g) == (1, 2) and cmsm.is_on_curve(g)
    assert cmsm.point_to_ints(cmsm.add(g, g)) == (G2X, G2Y)
    assert cmsm.point_to_ints(cmsm.scalar_mul(g, 2)) == (G2X, G2Y)
    assert cmsm.point_to_ints(cmsm.scalar_mul(g, pr.Q - 1)) == pr.neg(pr.G1)
    assert not cmsm.add(g, cmsm.neg(g)).any()  # infinity is (0, 0)
    assert np.array_equal(cmsm.add(g, np.zeros(8, dtype=np.uint64)), g)
    # field constants of the oracle and of the product headers: R = 2^256 mod p, R^2, -p^-1 mod 2^32
    hdr = open(os.path.join(EC_DIR, "field.cuh")).read()
    for name, mod in (("FpMod", pr.P), ("FrMod", pr.Q)):
        body = hdr[hdr.index("struct " + name):]
        body = body[:body.index("};")]
        vals = {k: int(v, 16) for k, v in re.findall(r"\b([A-Z]+\d?)\s*=\s*(0x[0-9a-fA-F]+)u", body)}
        limbs = lambda pre: sum(vals["%s%d" % (pre, i)] << (32 * i) for i in range(8))
        assert limbs("M") == mod
        assert limbs("R") == (1 << 256) % mod
        assert limbs("RR") == (1 << 512) % mod
        assert vals["NINV"] == (-pow(mod, -1, 1 << 32)) % (1 << 32)


def test_c_oracle_agrees_with_python_oracle(cmsm):
    import pyref_msm as pr
    rng = random.Random(7)
    n = 24
    pts = cmsm.gen_points(n)
    ints = [cmsm.point_to_ints(p) for p in pts]
    a, b = 0x1234567, 0x9E3779B97F4A7C15
    assert ints[0] == pr.mul(a, pr.G1) and ints[5] == pr.mul(a + 5 * b, pr.G1)
    assert all(pr.is_on_curve(p) for p in ints)
    sc = [0, 1, pr.Q - 1, 2, (1 << 253) + 5] + [rng.randrange(pr.Q) for _ in range(n - 5)]
    want = pr.multi_exp(ints, sc)
    assert cmsm.point_to_ints(cmsm.multiexp(pts, cmsm.scalars_regular(sc))) == want
    assert cmsm.point_to_ints(cmsm.multiexp(pts, cmsm.scalars_mont(sc), mont=True)) == want
    for nt in (1, 3):
        assert cmsm.point_to_ints(cmsm.multiexp(pts, cmsm.scalars_regular(sc), nthreads=nt)) == want
    # RawBytes, DeriveRandomnessFromPoint, InitialRandomnessHint
    for p_i, arr in zip(ints[:4], pts[:4]):
        assert cmsm.raw_bytes(arr) == pr.raw_bytes(p_i)
        assert cmsm.unlimbs(cmsm.derive_randomness_from_point(arr)) == pr.derive_randomness_from_point(p_i)
    zero = np.zeros(8, dtype=np.uint64)
    assert cmsm.raw_bytes(zero) == pr.raw_bytes(pr.INF) == bytes([0x40]) + bytes(63)
    assert cmsm.unlimbs(cmsm.derive_randomness_from_point(zero)) == pr.derive_randomness_from_point(pr.INF)
    kp, rnd = cmsm.initial_randomness(pts[:10], cmsm.scalars_regular(sc[:10]), pts[10:], cmsm.scalars_regular(sc[10:]))
    kp_py, rnd_py = pr.initial_randomness(ints[:10], sc[:10], ints[10:], sc[10:])
    assert cmsm.point_to_ints(kp) == kp_py and cmsm.unlimbs(rnd) == rnd_py


def test_bucket_method_oracle_agrees_with_the_naive_oracle(cmsm):
    """oracle/msm_oracle.c orc_g1_multiexp_buckets (gnark-crypto's published MultiExp algorithm restated: the CPU baseline and the
    second oracle the GPU tests use at 2^20 / 2^22) against one double-and-add per point"""
    rng = random.Random(21)
    for n in (1, 2, 3, 17, 100, 1000, 20000):
        pts = cmsm.gen_points(n, a=rng.randrange(cmsm.Q), b=rng.randrange(cmsm.Q))
        if n >= 17:
            pts[1] = 0
            pts[3] = pts[2]
            pts[5] = cmsm.neg(pts[4])
        sets = [[rng.randrange(cmsm.Q) for _ in range(n)], [rng.choice([0, 1, cmsm.Q - 1, 1 << 253, (1 << 253) - 1]) for _ in range(n)],
                [rng.randrange(cmsm.Q)] * n, [rng.randrange(1 << 40) for _ in range(n)]]
        for vals in (sets if n <= 1000 else sets[:1]):
            want = cmsm.multiexp(pts, cmsm.scalars_regular(vals))
            assert np.array_equal(cmsm.multiexp_buckets(pts, cmsm.scalars_regular(vals)), want), n
            assert np.array_equal(cmsm.multiexp_buckets(pts, cmsm.scalars_mont(vals), mont=True, nthreads=3), want), n
    assert not cmsm.multiexp_buckets(np.zeros((0, 8), dtype=np.uint64), np.zeros((0, 4), dtype=np.uint64)).any()


# ------------------------------------------------------------------------------------------------ 2. the kernel bodies on the host
def _edge(mod):
    r = (1 << 256) % mod
    return [0, 1, 2, mod - 1, mod - 2, r, mod - r, (1 << 253) - 1, (1 << 253), mod >> 1, (mod >> 1) + 1, (1 << 64) - 1, 1 << 64, (1 << 224) - 1,
            mod - (1 << 32), 0xFFFFFFFF, 1 << 32, (1 << 128) - 1]


def test_emulated_field_ops_match_big_integers(cmsm, emu):
    """f_mul / f_sqr / f_add / f_sub / f_inv / Montgomery conversions of csrc/ec/field.cuh for both moduli (Fp: G1 coordinates;
    Fr: the scalars' FromMont of hints.go:171), every pair of edge values and random ones"""
    rng = random.Random(11)
    for field, mod in ((0, cmsm.P), (1, cmsm.Q)):
        r = (1 << 256) % mod
        rinv = pow(r, -1, mod)
        e = _edge(mod)
        xs = [x for x in e for _ in e] + [rng.randrange(mod) for _ in range(600)]
        ys = [y for _ in e for y in e] + [rng.randrange(mod) for _ in range(600)]
        a = np.array([cmsm.limbs(v) for v in xs], dtype=np.uint64)
        b = np.array([cmsm.limbs(v) for v in ys], dtype=np.uint64)
        out = np.zeros_like(a)
        want = {
            0: lambda x, y: x * y * rinv % mod, 7: lambda x, y: x * y * rinv % mod, 1: lambda x, y: x * x * rinv % mod,
            2: lambda x, y: (x + y) % mod, 3: lambda x, y: (x - y) % mod, 5: lambda x, y: x * rinv % mod, 6: lambda x, y: x * r % mod,
        }
        for op, f in want.items():
            emu.emu_field_op(field, op, _p(a), _p(b), ctypes.c_size_t(len(xs)), _p(out))
            got = [cmsm.unlimbs(row) for row in out]
            assert got == [f(x, y) for x, y in zip(xs, ys)], (field, op)
        k = 60  # inversions are 380 products each
        emu.emu_field_op(field, 4, _p(a[-k:]), None, ctypes.c_size_t(k), _p(out[-k:]))
        assert [cmsm.unlimbs(row) for row in out[-k:]] == [pow(x * rinv % mod, -1, mod) * r % mod if x else 0 for x in xs[-k:]]
        ee = np.array([cmsm.limbs(v) for v in e], dtype=np.uint64)
        oo = np.zeros_like(ee)
        emu.emu_field_op(field, 4, _p(ee), None, ctypes.c_size_t(len(e)), _p(oo))
        assert [cmsm.unlimbs(row) for row in oo] == [pow(x * rinv % mod, -1, mod) * r % mod if x else 0 for x in e]


def test_emulated_group_law_handles_every_exceptional_case(cmsm, emu):
    """madd-2008-s / add-2008-s / dbl-2008-s-1 of csrc/ec/curve.cuh on G1 against the oracle's affine law: generic, P + P, P + (-P),
    infinity on either side, on trivial (ZZ = 1) and non-trivial representations, inlined and out-of-line multiplier"""
    pts = cmsm.gen_points(6)
    zero = np.zeros(8, dtype=np.uint64)
    cases = [(pts[0], pts[1]), (pts[2], pts[2]), (pts[3], cmsm.neg(pts[3])), (zero, pts[4]), (pts[4], zero), (zero, zero),
             (cmsm.generator(), cmsm.generator()), (cmsm.generator(), cmsm.neg(cmsm.generator()))]
    out = np.zeros(8, dtype=np.uint64)
    for a, b in cases:
        want = cmsm.add(a, b)
        for op in (0, 1, 4, 5):
            if op in (4, 5) and (not a.any() or (op == 5 and not b.any())):
                continue  # those build the operand from doublings of a non-zero point
            emu.emu_g1_op(op, _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out))
            assert np.array_equal(out, want), (op, cmsm.point_to_ints(a), cmsm.point_to_ints(b))
            assert cmsm.is_on_curve(out)
    for a in (pts[0], cmsm.generator(), zero):
        emu.emu_g1_op(2, _p(np.ascontiguousarray(a)), _p(zero), _p(out))
        assert np.array_equal(out, cmsm.add(a, a))
        for k in (0, 1, 2, 3, 16, 255, 256, 32767, 0xFFFFFFFF):
            kk = np.array([k, 0, 0, 0, 0, 0, 0, 0], dtype=np.uint64)
            emu.emu_g1_op(3, _p(np.ascontiguousarray(a)), _p(kk), _p(out))
            assert np.array_equal(out, cmsm.scalar_mul(a, k)), k
    # G1Affine.Add as the library's one-thread kernel runs it (result in Montgomery and regular form)
    out16 = np.zeros(16, dtype=np.uint64)
    for a, b in cases:
        emu.emu_g1_add_affine(_p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out16))
        want = cmsm.add(a, b)
        assert np.array_equal(out16[:8], want)
        x, y = cmsm.point_to_ints(want)
        assert cmsm.unlimbs(out16[8:12]) == x and cmsm.unlimbs(out16[12:16]) == y


def _scalar_sets(cmsm, n, rng):
    q = cmsm.Q
    yield "random", [rng.randrange(q) for _ in range(n)]
    yield "small", [rng.randrange(1 << 16) for _ in range(n)]
    yield "edge", [[0, 1, q - 1, q - 2, 2, (1 << 253), (1 << 253) - 1, ((1 << 254) - 1) % q, q >> 1][i % 9] % q for i in range(n)]
    yield "all equal", [0x1F3C5A7799BBDDFF0123456789ABCDEF0FEDCBA987654321 % q] * n  # the reference benchmark's shape: one value hashed 2^k times
    yield "all zero", [0] * n
    yield "digit boundaries", [sum(((1 << (c - 1)) + (i & 1)) << (c * w) for w in range(0, 254 // c)) % q for i, c in zip(range(n), [2, 3, 5, 8, 13, 16] * n)]


@pytest.mark.parametrize("n", [1, 2, 7, 64, 300])
def test_emulated_multiexp_matches_oracle(cmsm, emu, n):
    """the whole launch sequence of msm_enqueue on the host executor == sum_i s_i P_i of the oracle, for every window width the
    plan can pick, task sizes that split buckets, both execution orders, Montgomery and regular scalars"""
    rng = random.Random(100 + n)
    pts = cmsm.gen_points(n, a=rng.randrange(cmsm.Q), b=rng.randrange(cmsm.Q))
    if n >= 7:  # the caller's bases are arbitrary: infinity, repeated and opposite points among them
        pts[1] = 0
        pts[3] = pts[2]
        pts[5] = cmsm.neg(pts[4])
    for name, vals in _scalar_sets(cmsm, n, rng):
        reg, mont = cmsm.scalars_regular(vals), cmsm.scalars_mont(vals)
        want = cmsm.multiexp(pts, reg)
        plans = [(0, 0), (2, 0), (3, 1), (7, 2), (16, 0)] if n <= 64 else [(0, 0), (4, 3), (11, 0)]
        for c, T in plans:
            for rev in (0, 1):
                rc, out, plan = _emu_msm(emu, pts, reg, 0, c, T, rev)
                assert rc == 0 and np.array_equal(out[:8], want), (name, c, T, rev)
                x, y = cmsm.point_to_ints(want)
                assert cmsm.unlimbs(out[8:12]) == x and cmsm.unlimbs(out[12:16]) == y
        rc, out, _ = _emu_msm(emu, pts, mont, 1)
        assert rc == 0 and np.array_equal(out[:8], want), name


def test_emulated_multiexp_larger_and_skewed(cmsm, emu):
    """4 096 points: the cost model's own plan, a forced 16-bit window, and all scalars equal with small tasks (one bucket per window
    holds every point and is cut into 256 tasks)"""
    n = 4096
    rng = random.Random(9)
    pts = cmsm.gen_points(n)
    sc = cmsm.scalars_regular([rng.randrange(cmsm.Q) for _ in range(n)])
    want = cmsm.multiexp(pts, sc)
    rc, out, plan = _emu_msm(emu, pts, sc)
    assert rc == 0 and np.array_equal(out[:8], want) and 2 <= plan[0] <= 16
    rc, out, _ = _emu_msm(emu, pts, sc, c=16)
    assert rc == 0 and np.array_equal(out[:8], want)
    same = cmsm.scalars_regular([0x2545F4914F6CDD1D2545F4914F6CDD1D2545F4914F6CDD1D % cmsm.Q] * n)
    want = cmsm.multiexp(pts, same)
    rc, out, _ = _emu_msm(emu, pts, same, T=16, reverse=1)
    assert rc == 0 and np.array_equal(out[:8], want)
    # task size 1: every hot bucket has 4 096 partial sums = 128 reduction groups of 32 (the second level of the bucket reduction)
    rc, out, _ = _emu_msm(emu, pts, same, T=1)
    assert rc == 0 and np.array_equal(out[:8], want)
    three = cmsm.scalars_regular([(i % 3 + 5) * 0x123456789ABCDEF123456789 % cmsm.Q for i in range(n)])  # key / msg / out repeated: the hint's shape
    rc, out, _ = _emu_msm(emu, pts, three, T=3, reverse=1)
    assert rc == 0 and np.array_equal(out[:8], cmsm.multiexp(pts, three))


def test_emulated_multiexp_rejects_unreduced_scalars(cmsm, emu):
    pts = cmsm.gen_points(3)
    sc = cmsm.scalars_regular([1, 2, 3])
    sc[1] = np.array(cmsm.limbs(cmsm.Q), dtype=np.uint64)  # q itself: fr.Element never holds it
    rc, _, _ = _emu_msm(emu, pts, sc)
    assert rc & 1
    # in Montgomery form every 256-bit pattern below 2^256 reduces (FromMont), no error
    sc_m = cmsm.scalars_mont([1, 2, 3])
    rc, out, _ = _emu_msm(emu, pts, sc_m, mont=1)
    assert rc == 0 and np.array_equal(out[:8], cmsm.multiexp(pts, cmsm.scalars_regular([1, 2, 3])))


def test_plan_invariants(emu):
    """W * c >= 255 (the top signed digit never carries out), 2^(c-1) buckets, chunks cover the buckets, the task bound holds"""
    plan = np.zeros(8, dtype=np.uint32)
    dummy_p, dummy_s = np.zeros((1, 8), dtype=np.uint64), np.zeros((1, 4), dtype=np.uint64)
    out = np.zeros(16, dtype=np.uint64)
    for c in range(2, 17):
        emu.emu_msm(_p(dummy_p), _p(dummy_s), ctypes.c_size_t(1), 0, c, 0, 0, _p(out), _p(plan))
        cc, W, T, L, nch = (int(v) for v in plan[:5])
        assert cc == c and W * c >= 255 and (W - 1) * c <= 254 and T >= 1
        B = 1 << (c - 1)
        assert L * nch >= B and L * (nch - 1) < B
        assert not out.any()  # 0 * P = infinity


def test_emulated_multiexp_randomised_sweep(cmsm, emu):
    """80 random cases: size, window width, task size, scalar shape (incl. values on every signed-digit boundary of the chosen width),
    scalar form, launch order, bases drawn with repetition and with infinity / negated points mixed in"""
    rng = random.Random(2026)
    q = cmsm.Q
    pool = cmsm.gen_points(300, a=rng.randrange(q), b=rng.randrange(q))
    for case in range(80):
        n = rng.randrange(1, rng.choice([2, 3, 5, 17, 33, 100, 257]) + 1)
        c = rng.choice([0, 0, 2, 3, 4, 5, 6, 8, 9, 12, 13, 15, 16])
        T = rng.choice([0, 0, 1, 2, 3, 5, 16, 100])
        kind = rng.randrange(6)
        if kind == 0:
            vals = [rng.randrange(q) for _ in range(n)]
        elif kind == 1:
            vals = [rng.randrange(1 << rng.choice([1, 8, 31, 32, 33, 64, 128, 253])) for _ in range(n)]
        elif kind == 2:
            vals = [rng.choice([0, 1, q - 1, q - 2, 1 << 253, (1 << 253) - 1, q >> 1]) for _ in range(n)]
        elif kind == 3:
            vals = [rng.randrange(q)] * n
        elif kind == 4:
            cc = max(c, 2)
            vals = [((1 << (cc - 1)) * sum(1 << (cc * w) for w in range(254 // cc)) + rng.randrange(3) - 1) % q for _ in range(n)]
        else:
            vals = [(q - rng.randrange(1 << 20)) % q for _ in range(n)]
        pts = pool[[rng.randrange(300) for _ in range(n)]].copy()
        for j in range(n):
            r = rng.random()
            if r < 0.05:
                pts[j] = 0
            elif r < 0.15:
                pts[j] = cmsm.neg(pts[j])
        mont = int(rng.random() < 0.3)
        sc = (cmsm.scalars_mont if mont else cmsm.scalars_regular)(vals)
        rc, out, _ = _emu_msm(emu, pts, sc, mont, c, T, rng.randrange(2))
        assert rc == 0 and np.array_equal(out[:8], cmsm.multiexp(pts, cmsm.scalars_regular(vals))), (case, n, c, T, kind, mont)


# ------------------------------------------------------------------------------------------------ 2b. G2 (prove.go:277, Bs)
def test_oracle_g2_known_answers(cmsm):
    """the published generator of G2 (EIP-197 / gnark-crypto) lies on y^2 = x^3 + 3/(9+u) and has order q; C oracle == Python"""
    import pyref_msm as pr
    assert pr.g2_is_on_curve(pr.G2) and pr.g2_mul(pr.Q, pr.G2) == pr.INF2 and pr.g2_mul(pr.Q - 1, pr.G2) == pr.g2_neg(pr.G2)
    g = cmsm.g2_generator()
    assert cmsm.g2_point_to_ints(g) == pr.G2 and cmsm.g2_is_on_curve(g)
    bad = g.copy()
    bad[0] ^= np.uint64(1)
    assert not cmsm.g2_is_on_curve(bad)
    assert cmsm.g2_point_to_ints(cmsm.g2_add(g, g)) == pr.g2_add(pr.G2, pr.G2)
    assert cmsm.g2_point_to_ints(cmsm.g2_scalar_mul(g, pr.Q - 1)) == pr.g2_neg(pr.G2)
    assert not cmsm.g2_add(g, cmsm.g2_neg(g)).any()
    rng = random.Random(3)
    n = 10
    pts = cmsm.g2_gen_points(n)
    ints = [cmsm.g2_point_to_ints(p) for p in pts]
    assert ints[3] == pr.g2_mul(0x7654321 + 3 * 0xD1B54A32D192ED03, pr.G2) and all(pr.g2_is_on_curve(p) for p in ints)
    sc = [0, 1, pr.Q - 1] + [rng.randrange(pr.Q) for _ in range(n - 3)]
    want = pr.g2_multi_exp(ints, sc)
    assert cmsm.g2_point_to_ints(cmsm.g2_multiexp(pts, cmsm.scalars_regular(sc))) == want
    assert cmsm.g2_point_to_ints(cmsm.g2_multiexp(pts, cmsm.scalars_mont(sc), mont=True)) == want


def test_emulated_fp2_ops_match_big_integers(cmsm, emu):
    """fptower.E2 as the kernels compute it (Karatsuba product, complex squaring, norm inversion) against the schoolbook definition"""
    import pyref_msm as pr
    P, RP, RI = cmsm.P, cmsm.RP, cmsm.RP_INV
    rng = random.Random(5)
    e = [0, 1, P - 1, P - 2, RP, (1 << 253) - 1, P >> 1]
    vals = [(x, y) for x in e for y in e] + [(rng.randrange(P), rng.randrange(P)) for _ in range(300)]
    vb = list(reversed(vals))
    tom = lambda vs: np.array([cmsm.limbs(v[0] * RP % P) + cmsm.limbs(v[1] * RP % P) for v in vs], dtype=np.uint64)
    frm = lambda arr: [(cmsm.unlimbs(r[:4]) * RI % P, cmsm.unlimbs(r[4:]) * RI % P) for r in arr]
    a, b = tom(vals), tom(vb)
    out = np.zeros_like(a)
    ops = {0: pr.f2_mul, 5: pr.f2_mul, 1: lambda x, y: pr.f2_mul(x, x), 2: pr.f2_add, 3: pr.f2_sub,
           4: lambda x, y: pr.f2_inv(x) if x != (0, 0) else (0, 0)}
    for op, f in ops.items():
        k = len(vals) if op != 4 else 80
        emu.emu_fp2_op(op, _p(a), _p(b), ctypes.c_size_t(k), _p(out))
        assert frm(out[:k]) == [f(x, y) for x, y in zip(vals[:k], vb[:k])], op


def test_emulated_g2_group_law_handles_every_exceptional_case(cmsm, emu):
    pts = cmsm.g2_gen_points(6)
    zero = np.zeros(16, dtype=np.uint64)
    g = cmsm.g2_generator()
    cases = [(pts[0], pts[1]), (pts[2], pts[2]), (pts[3], cmsm.g2_neg(pts[3])), (zero, pts[4]), (pts[4], zero), (zero, zero), (g, g), (g, cmsm.g2_neg(g))]
    out = np.zeros(16, dtype=np.uint64)
    out32 = np.zeros(32, dtype=np.uint64)
    for a, b in cases:
        want = cmsm.g2_add(a, b)
        for op in (0, 1, 4, 5):
            if op in (4, 5) and (not a.any() or (op == 5 and not b.any())):
                continue
            emu.emu_g2_op(op, _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out))
            assert np.array_equal(out, want), op
            assert cmsm.g2_is_on_curve(out)
        emu.emu_g2_add_affine(_p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out32))
        assert np.array_equal(out32[:16], want)
        (x0, x1), (y0, y1) = cmsm.g2_point_to_ints(want)
        assert [cmsm.unlimbs(out32[16 + 4 * k:20 + 4 * k]) for k in range(4)] == [x0, x1, y0, y1]  # the regular-form copy
    for a in (pts[0], g, zero):
        emu.emu_g2_op(2, _p(np.ascontiguousarray(a)), _p(zero), _p(out))
        assert np.array_equal(out, cmsm.g2_add(a, a))
        for k in (0, 1, 2, 3, 255, 256, 0xFFFFFFFF):
            kk = np.zeros(16, dtype=np.uint64)
            kk[0] = k
            emu.emu_g2_op(3, _p(np.ascontiguousarray(a)), _p(kk), _p(out))
            assert np.array_equal(out, cmsm.g2_scalar_mul(a, k)), k


@pytest.mark.parametrize("n", [1, 2, 7, 64, 300])
def test_emulated_g2_multiexp_matches_oracle(cmsm, emu, n):
    """the same launch sequence over the quadratic extension: G2Affine.MultiExp"""
    rng = random.Random(300 + n)
    pts = cmsm.g2_gen_points(n, a=rng.randrange(cmsm.Q), b=rng.randrange(cmsm.Q))
    if n >= 7:
        pts[1] = 0
        pts[3] = pts[2]
        pts[5] = cmsm.g2_neg(pts[4])
    for name, vals in _scalar_sets(cmsm, n, rng):
        if n > 64 and name in ("small", "digit boundaries"):
            continue
        reg = cmsm.scalars_regular(vals)
        want = cmsm.g2_multiexp(pts, reg)
        for c, T in ([(0, 0), (2, 0), (3, 1), (16, 0)] if n <= 2 else [(0, 0), (2, 0), (3, 1)] if n <= 64 else [(0, 0), (5, 3)]):  # 2^15 Fp2 buckets per window are slow on the CPU
            for rev in (0, 1):
                out = np.zeros(32, dtype=np.uint64)
                rc = emu.emu_g2_msm(_p(pts), _p(reg), ctypes.c_size_t(n), 0, c, T, rev, _p(out), None)
                assert rc == 0 and np.array_equal(out[:16], want), (name, c, T, rev)
        out = np.zeros(32, dtype=np.uint64)
        rc = emu.emu_g2_msm(_p(pts), _p(cmsm.scalars_mont(vals)), ctypes.c_size_t(n), 1, 0, 0, 0, _p(out), None)
        assert rc == 0 and np.array_equal(out[:16], want), name


# ------------------------------------------------------------------------------------------------ 3. the library without a GPU
def _has_gpu():
    try:
        return subprocess.run(["nvidia-smi", "-L"], capture_output=True, timeout=20).returncode == 0
    except Exception:
        return False


def test_ec_library_exports_every_declared_symbol():
    from gkrb200 import ec
    hdr = open(os.path.join(ROOT, "include", "gkrb200_ec.h")).read()
    names = sorted(set(re.findall(r"\b(gkrb200ec_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 14
    L = ec.lib()
    for n in names:
        assert hasattr(L, n), "symbol %s declared in include/gkrb200_ec.h is not exported" % n
    out = subprocess.run(["nm", "-D", "--defined-only", ec.SO_PATH], capture_output=True, text=True).stdout
    assert set(names) <= set(re.findall(r" T (gkrb200ec_[a-z0-9_]+)", out))
    assert L.gkrb200ec_version().startswith(b"gkrb200ec")
    archs = set(re.findall(r"sm_(\d+a?)", subprocess.run(["cuobjdump", "--list-elf", ec.SO_PATH], capture_output=True, text=True).stdout))
    assert archs == {"100a"}, archs
    ldd = subprocess.run(["ldd", ec.SO_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "emu" not in ldd
    for f in os.listdir(EC_DIR) + ["../../gkrb200/ec.py"]:
        txt = open(os.path.join(EC_DIR, f), errors="replace").read()
        assert "cmsm" not in txt and "pyref" not in txt and "msm_oracle" not in txt and "libmsmoracle" not in txt, f


def test_ec_header_is_valid_c99(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "gkrb200_ec.h"\nint main(void) { gkrb200ec_stats s; (void)s; return GKRB200EC_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-c", str(src), "-o",
                           str(tmp_path / "t.o")])


def test_plain_c_program_reproduces_keccak_kat_through_the_ec_library(tmp_path):
    """the header and the library from C99, no Python in between: legacy Keccak-256("abc") and RawBytes of the generator (1, 2)"""
    from gkrb200 import ec
    src = tmp_path / "kat.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "gkrb200_ec.h"
int main(void) {
    uint8_t h[32], raw[64];
    if (gkrb200ec_keccak256((const uint8_t *)"abc", 3, h) != GKRB200EC_OK) return 1;
    for (int i = 0; i < 32; i++) printf("%02x", h[i]);
    printf("\n");
    /* (1, 2) in Montgomery form: R mod p, 2R mod p */
    uint64_t g[8] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL,
                     0xa6ba871b8b1e1b3aULL, 0x14f1d651eb8e167bULL, 0xccdd46def0f28c58ULL, 0x1c14ef83340fbe5eULL};
    if (gkrb200ec_g1_raw_bytes(g, raw) != GKRB200EC_OK) return 2;
    for (int i = 0; i < 64; i++) printf("%02x", raw[i]);
    printf("\n");
    if (gkrb200ec_g1_raw_bytes(NULL, raw) != GKRB200EC_ERR_ARG) return 3;
    printf("%s\n", gkrb200ec_last_error());
    return 0;
}
''')
    exe = tmp_path / "kat"
    lib_dir = os.path.dirname(ec.SO_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L" + lib_dir,
                           "-lgkrb200ec", "-Wl,-rpath," + lib_dir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.split("\n")
    assert lines[0] == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    assert lines[1] == "%064x%064x" % (1, 2)
    assert "null" in lines[2]


def test_ec_host_entry_points_match_oracle(cmsm):
    """RawBytes, legacy Keccak-256 and DeriveRandomnessFromPoint (hints.go:147-159) run on the host inside the product"""
    from gkrb200 import ec
    for data, want in KECCAK_KAT:
        assert ec.LegacyKeccak256(data).hex() == want
    rng = random.Random(3)
    for n in [1, 55, 135, 136, 137, 272, 273, 1000]:
        data = bytes(rng.randrange(256) for _ in range(n))
        assert ec.LegacyKeccak256(data) == cmsm.keccak256(data), n
    pts = list(cmsm.gen_points(8)) + [np.zeros(8, dtype=np.uint64), cmsm.generator()]
    for pt in pts:
        assert ec.RawBytes(pt) == cmsm.raw_bytes(pt)
        assert np.array_equal(ec.DeriveRandomnessFromPoint(pt), cmsm.derive_randomness_from_point(pt))
    assert ec.RawBytes(cmsm.generator()) == (1).to_bytes(32, "big") + (2).to_bytes(32, "big")


def test_ec_no_cpu_fallback_without_a_device():
    from gkrb200 import ec
    L = ec.lib()
    h = ctypes.c_void_p()
    assert L.gkrb200ec_init(None, 0, None) == -1
    assert L.gkrb200ec_set_plan(None, 0, 0) == -1 and b"null" in L.gkrb200ec_last_error()
    if _has_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(ec.GkrB200EcError) as e:
        ec.EcContext(device=0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    assert L.gkrb200ec_init(ctypes.byref(h), 0, None) == -2 and not h.value
