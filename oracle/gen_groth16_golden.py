"""Writes tests/golden/groth16_side.json: small input/output vectors of the Groth16-side operations computed by
the ORACLE (oracle/pyref_msm.py, oracle/pyref_fft.py: Python big integers).  They are regression pins for the oracles, the emulated
kernels and the device library -- and a ready-made cross-check for a maintainer with a Go toolchain: INTEGRATION.md section 7 shows the
few lines of Go that feed the same inputs to gnark-crypto (G1Affine.MultiExp, fft.Domain, computeH, DeriveRandomnessFromPoint).
NOT reference outputs: the reference cannot be run in this image (no Go); values are decimal strings of REGULAR-form integers.

    python oracle/gen_groth16_golden.py
"""
import json
import os
import random

import pyref_fft as pf
import pyref_msm as pr

rng = random.Random(0x6b6b72)
out = {"note": "computed by oracle/pyref_msm.py and oracle/pyref_fft.py (Python big integers), not by the Go reference; regular-form decimal strings",
       "curve": "BN254: G1 y^2 = x^3 + 3 over Fp, generator (1, 2); scalar field Fr"}

# G1 multi-exponentiation: points k_i * G, scalars incl. edge values
ks = [1, 2, 3, 0x1234567, pr.Q - 1, 5, 5, 7] + [rng.randrange(1, pr.Q) for _ in range(8)]
pts = [pr.mul(k, pr.G1) for k in ks]
pts[6] = pr.neg(pts[5])  # a pair of opposite points
pts[7] = pr.INF          # a point at infinity among the bases
scalars = [0, 1, pr.Q - 1, 2, (1 << 253) + 5, 11, 11, 13] + [rng.randrange(pr.Q) for _ in range(8)]
res = pr.multi_exp(pts, scalars)
out["g1_multiexp"] = {"points": [[str(x), str(y)] for x, y in pts], "scalars": [str(s) for s in scalars], "result": [str(res[0]), str(res[1])]}

# DeriveRandomnessFromPoint / InitialRandomnessHint (hints.go:147-192)
out["derive_randomness"] = [{"point": [str(p[0]), str(p[1])], "raw_bytes_hex": pr.raw_bytes(p).hex(), "keccak256_hex": pr.keccak256(pr.raw_bytes(p)).hex(),
                             "randomness": str(pr.derive_randomness_from_point(p))} for p in (pr.G1, res, pr.INF)]
krs_priv, rnd = pr.initial_randomness(pts[:8], scalars[:8], pts[8:], scalars[8:])
out["initial_randomness_hint"] = {"pub": "points/scalars 0..7 of g1_multiexp", "priv": "points/scalars 8..15", "krs_gkr_priv": [str(krs_priv[0]), str(krs_priv[1])],
                                  "initial_randomness": str(rnd)}

# G2 multi-exponentiation
ks2 = [1, 2, pr.Q - 1] + [rng.randrange(1, pr.Q) for _ in range(3)]
pts2 = [pr.g2_mul(k, pr.G2) for k in ks2]
sc2 = [3, pr.Q - 2, 1 << 200] + [rng.randrange(pr.Q) for _ in range(3)]
res2 = pr.g2_multi_exp(pts2, sc2)
f2s = lambda e: [str(e[0]), str(e[1])]
out["g2_multiexp"] = {"points": [[f2s(p[0]), f2s(p[1])] for p in pts2], "scalars": [str(s) for s in sc2], "result": [f2s(res2[0]), f2s(res2[1])],
                      "layout": "[[X.A0, X.A1], [Y.A0, Y.A1]]"}

# fft.NewDomain(m, 1, true) and computeH (prove.go:310-366) on 11 constraints (cardinality 16): arbitrary a, b, c and a satisfied system
dom = pf.Domain(11)
a = [rng.randrange(pf.Q) for _ in range(11)]
b = [rng.randrange(pf.Q) for _ in range(11)]
c = [rng.randrange(pf.Q) for _ in range(11)]
c_sat = [x * y % pf.Q for x, y in zip(a, b)]
out["domain_11"] = {"cardinality": dom.n, "generator": str(dom.generator), "finer_generator": str(dom.finer_generator), "cardinality_inv": str(dom.cardinality_inv)}
v = [rng.randrange(pf.Q) for _ in range(16)]
out["fft_16"] = {"input": [str(x) for x in v],
                 "fft_dif_coset0": [str(x) for x in pf.fft(dom, v, pf.DIF, 0)], "fft_dit_coset1": [str(x) for x in pf.fft(dom, v, pf.DIT, 1)],
                 "fftinverse_dif_coset1": [str(x) for x in pf.fft_inverse(dom, v, pf.DIF, 1)], "fftinverse_dit_coset0": [str(x) for x in pf.fft_inverse(dom, v, pf.DIT, 0)]}
out["compute_h_11"] = {"a": [str(x) for x in a], "b": [str(x) for x in b], "c": [str(x) for x in c], "h": [str(x) for x in pf.compute_h(a, b, c, dom)],
                       "c_satisfied": [str(x) for x in c_sat], "h_satisfied": [str(x) for x in pf.compute_h(a, b, c_sat, dom)],
                       "h_satisfied_by_long_division": [str(x) for x in pf.compute_h_by_division(a, b, c_sat, dom)]}

dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "groth16_side.json")
json.dump(out, open(dst, "w"), indent=1)
print("wrote", dst, os.path.getsize(dst), "bytes")
