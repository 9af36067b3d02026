"""First measurement of the Groth16-side library (libgkrb200ec.so) on one B200 -- NOT yet run (it was written after
the round's GPU budget was spent).  One JSON line per operation, bench.py-style: CUDA-event time of the operation's kernels (the
library times them on its own stream), algorithmic work against the integer-multiply peak, and the oracle on the host cores as a
labelled baseline.

    python tools/gpu_groth16_side.py [log2_n = 20] [reps = 5]

Workloads: the InitialRandomnessHint shape (hints.go:162-192: 3 * 2^k GKR inputs/outputs as scalars; here one G1 multi-exponentiation
of 2^k points), the G2 multi-exponentiation of prove.go:277, computeH (prove.go:310-366) on 2^k constraints.  Every result is checked:
closed form on bases with known discrete logs (multi-exponentiations), the quotient identity at a random point (computeH).
The CPU baselines are the ORACLE's restatements (the bucket method of gnark-crypto's MultiExp, one thread per window; the textbook
transform, single thread) in portable C: slower than gnark-crypto's assembly, printed with that label and on a bounded sample.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gkr-mimc_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

import cfft
import cmsm
from gkrb200 import ec

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = 1 << lg
MAC_PEAK = 8.1e12        # carry-chained IMAD.WIDE.U32 per second measured on this pool's B200
MUL_PEAK = 66.4e9        # the library's Montgomery product in isolation (136 wide multiply-adds each)
MADD_MACS = 8 * 136 + 2 * 108   # madd-2008-s: 8 products + 2 squarings


def rand252(rng, k):
    w = rng.integers(0, 1 << 63, size=(k, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(k, 4), dtype=np.uint64)
    w[:, 3] >>= np.uint64(12)
    return w


def limb_sums(arr):
    idx = np.arange(arr.shape[0], dtype=np.uint64)
    tot = wtot = 0
    for j in range(4):
        for k in range(4):
            piece = (arr[:, j] >> np.uint64(16 * k)) & np.uint64(0xFFFF)
            tot += int(piece.sum()) << (64 * j + 16 * k)
            wtot += int((piece * idx).sum()) << (64 * j + 16 * k)
    return tot, wtot


def line(**kw):
    print(json.dumps(kw), flush=True)


cmsm.build()
cfft.build()
rng = np.random.default_rng(lg)
ctx = ec.EcContext(device=0)

# ---- G1 multi-exponentiation
a, b = 0x1234567, 0x9E3779B97F4A7C15
pts = cmsm.gen_points(n, a=a, b=b)
ctx.SetBases(0, pts)
s = rand252(rng, n)
best = None
for _ in range(reps):
    got = ctx.MultiExp(0, s)
    ms = ctx.stats().last_device_ms
    best = ms if best is None else min(best, ms)
tot, wtot = limb_sums(s)
ok = bool(np.array_equal(got, cmsm.scalar_mul(cmsm.generator(), (a * tot + b * wtot) % cmsm.Q)))
st = ctx.stats()
macs = n * st.last_windows * MADD_MACS
sample = min(n, 1 << 20)
t0 = time.time()
cpu_res = cmsm.multiexp_buckets(pts[:sample], s[:sample])
cpu_s = time.time() - t0
if sample == n:
    ok = ok and bool(np.array_equal(got, cpu_res))
line(op="G1Affine.MultiExp", n=n, window_bits=st.last_c, windows=st.last_windows, device_ms=best, points_per_s=n / (best * 1e-3), parity_closed_form=ok,
     roofline={"bound": "integer", "achieved": macs / (best * 1e-3) / 1e12, "peak": MAC_PEAK / 1e12, "unit": "T wide MAC/s", "frac": macs / (best * 1e-3) / MAC_PEAK,
               "algorithmic": "n * windows * (8 * 136 + 2 * 108) wide multiply-adds (mixed additions only)"},
     cpu_baseline={"kind": "port (oracle/msm_oracle.c orc_g1_multiexp_buckets: gnark-crypto's published bucket-method algorithm restated in C, one thread per window; "
                           "plain __int128 field arithmetic, not gnark-crypto's assembly)", "cores": min(cmsm.threads(), 16), "sample": "%d points" % sample,
                   "points_per_s": sample / cpu_s})

# ---- G2 multi-exponentiation
n2 = min(n, 1 << 20)
a2, b2 = 0x7654321, 0xD1B54A32D192ED03
pts2 = cmsm.g2_gen_points(n2, a=a2, b=b2)
ctx.SetBasesG2(1, pts2)
s2 = s[:n2]
best = None
for _ in range(reps):
    got = ctx.MultiExpG2(1, s2)
    ms = ctx.stats().last_device_ms
    best = ms if best is None else min(best, ms)
tot, wtot = limb_sums(s2)
ok = bool(np.array_equal(got, cmsm.g2_scalar_mul(cmsm.g2_generator(), (a2 * tot + b2 * wtot) % cmsm.Q)))
st = ctx.stats()
macs = n2 * st.last_windows * (8 * 3 * 136 + 2 * 2 * 136)  # Karatsuba product = 3 base products, complex squaring = 2
line(op="G2Affine.MultiExp", n=n2, window_bits=st.last_c, windows=st.last_windows, device_ms=best, points_per_s=n2 / (best * 1e-3), parity_closed_form=ok,
     roofline={"bound": "integer", "achieved": macs / (best * 1e-3) / 1e12, "peak": MAC_PEAK / 1e12, "unit": "T wide MAC/s", "frac": macs / (best * 1e-3) / MAC_PEAK})
ctx.SetBasesG2(1, np.zeros((0, 16), dtype=np.uint64))

# ---- computeH
assert ctx.NewDomain(n) == n
av, bv = rand252(rng, n), rand252(rng, n)
cv = cfft.mul_elementwise(av, bv)
best = None
for _ in range(reps):
    h = ctx.ComputeH(av, bv, cv)
    ms = ctx.stats().last_fft_device_ms
    best = ms if best is None else min(best, ms)
z = rand252(rng, 1)[0]
ok = bool(cfft.quotient_identity_holds(av, bv, cv, h, n, z))
muls = 7 * lg * (n // 2) + 3 * 2 * n + 2 * n + 3 * n  # butterflies + 3 fused scalings (2 products each) + pointwise (2) + last scaling with FromMont (3)
passes = 7 * ((lg + 2) // 3) + 3 + 1 + 1
sample = min(n, 1 << 18)
t0 = time.time()
cfft.compute_h(av[:sample], bv[:sample], cv[:sample], sample)
cpu_s = time.time() - t0
line(op="computeH", n=n, device_ms=best, constraints_per_s=n / (best * 1e-3), parity_quotient_identity=ok,
     roofline={"bound": "integer", "achieved": muls / (best * 1e-3) / 1e9, "peak": MUL_PEAK / 1e9, "unit": "G Fr products/s", "frac": muls / (best * 1e-3) / MUL_PEAK,
               "hbm_gbs": passes * 64 * n / (best * 1e-3) / 1e9, "passes_over_the_array": passes},
     cpu_baseline={"kind": "port (oracle: textbook transform, single thread)", "cores": 1, "sample": "2^%d constraints" % (sample.bit_length() - 1),
                   "constraints_per_s": sample / cpu_s})
ctx.close()
