"""Oracle composition of ComputeGroth16Proof (prover/gadget/prove.go:100-306) from the oracle's own pieces (cmsm: one double-and-add
per point; cfft: textbook transforms) -- TEST INFRASTRUCTURE ONLY.

    h   = computeH(a, b, c)                                                      prove.go:126, 310-366
    ar  = MultiExp(pk.G1.A, wireValuesA) + pk.G1.Alpha + r pk.G1.Delta           prove.go:199-210
    bs1 = MultiExp(pk.G1.B, wireValuesB) + pk.G1.Beta + s pk.G1.Delta            prove.go:186-196
    krs = -(r s) pk.G1.Delta + MultiExp(pk.G1.Z, h) + s ar + r bs1               prove.go:212-262 (the fork removed the pk.G1.K term)
    bs  = MultiExp(pk.G2.B, wireValuesB) + s pk.G2.Delta + pk.G2.Beta            prove.go:265-292
r, s are Python integers (the reference draws them at random, prove.go:154-161; a test fixes them).
"""
import numpy as np

import cfft
import cmsm


def synthetic_proving_key(n_a, n_b, cardinality, seed=1):
    """point arrays of the shapes groth16.ProvingKey holds, with known discrete logs (any points do: the prover never looks inside)"""
    import random
    rng = random.Random(seed)
    rq = lambda: rng.randrange(1, cmsm.Q)
    g1, g2 = cmsm.generator(), cmsm.g2_generator()
    return {
        "g1_a": cmsm.gen_points(n_a, a=rq(), b=rq()), "g1_b": cmsm.gen_points(n_b, a=rq(), b=rq()),
        "g1_z": cmsm.gen_points(cardinality, a=rq(), b=rq()), "g2_b": cmsm.g2_gen_points(n_b, a=rq(), b=rq()),
        "g1_alpha": cmsm.scalar_mul(g1, rq()), "g1_beta": cmsm.scalar_mul(g1, rq()), "g1_delta": cmsm.scalar_mul(g1, rq()),
        "g2_beta": cmsm.g2_scalar_mul(g2, rq()), "g2_delta": cmsm.g2_scalar_mul(g2, rq()),
    }


def compute_groth16_proof(pk, a, b, c, wire_values_a, wire_values_b, r, s, cardinality):
    """a, b, c: (n, 4) Montgomery; wire_values_*: (n, 4) REGULAR form -> (Ar (8,), Bs (16,), Krs (8,))"""
    q = cmsm.Q
    h = cfft.compute_h(a, b, c, cardinality)
    kr = (-(r * s)) % q
    d_r, d_s, d_kr = (cmsm.scalar_mul(pk["g1_delta"], k) for k in (r, s, kr))
    bs1 = cmsm.add(cmsm.add(cmsm.multiexp(pk["g1_b"], wire_values_b), pk["g1_beta"]), d_s)
    ar = cmsm.add(cmsm.add(cmsm.multiexp(pk["g1_a"], wire_values_a), pk["g1_alpha"]), d_r)
    krs = cmsm.add(d_kr, cmsm.multiexp(pk["g1_z"], h))
    krs = cmsm.add(krs, cmsm.scalar_mul(ar, s))
    krs = cmsm.add(krs, cmsm.scalar_mul(bs1, r))
    bs = cmsm.g2_add(cmsm.g2_add(cmsm.g2_multiexp(pk["g2_b"], wire_values_b), cmsm.g2_scalar_mul(pk["g2_delta"], s)), pk["g2_beta"])
    return ar, bs, krs


def fr_mont(v):
    return np.array(cmsm.limbs(v % cmsm.Q * cmsm.RQ % cmsm.Q), dtype=np.uint64)
