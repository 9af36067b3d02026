"""GPU parity test of ComputeGroth16Proof in one call (gkrb200ec_groth16_prove; prover/gadget/prove.go:100-306): computeH, the three
G1 multi-exponentiations, the G2 multi-exponentiation and the single-point operations on the device, against the oracle's own
composition of its own pieces (oracle/cgroth16.py).  r and s are fixed by the test (the reference draws them at random); with them
given the proof elements Ar, Bs, Krs are a deterministic function of the inputs and must match bit for bit.
(Sorts after the other Groth16-side GPU tests, which check the pieces.)"""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n_a,n_b,mont", [(5, 7, 6, 0), (100, 90, 77, 1), (1000, 1500, 1200, 0), (4096, 4000, 3000, 1)])
def test_compute_groth16_proof_matches_oracle(m, n_a, n_b, mont):
    import cfft
    import cgroth16
    import cmsm
    from gkrb200 import ec
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
    want = cgroth16.compute_groth16_proof(pk, a, b, c, cmsm.scalars_regular(wa_vals), cmsm.scalars_regular(wb_vals), r, s, n)
    conv = cmsm.scalars_mont if mont else cmsm.scalars_regular
    with ec.EcContext(device=0) as ctx:
        assert ctx.NewDomain(m) == n
        ctx.SetBases(0, pk["g1_a"])
        ctx.SetBases(1, pk["g1_b"])
        ctx.SetBases(2, pk["g1_z"])
        ctx.SetBasesG2(3, pk["g2_b"])
        got = ctx.ComputeGroth16Proof((0, 1, 2, 3), pk, a, b, c, conv(wa_vals), conv(wb_vals), cgroth16.fr_mont(r), cgroth16.fr_mont(s),
                                      ec.SCALARS_MONTGOMERY if mont else ec.SCALARS_REGULAR)
        for g, w, name in zip(got, want, ("Ar", "Bs", "Krs")):
            assert np.array_equal(g, w), name
        assert cmsm.is_on_curve(got[0]) and cmsm.g2_is_on_curve(got[1]) and cmsm.is_on_curve(got[2])
        # a second call with other randomness changes every element (r, s enter all three)
        got2 = ctx.ComputeGroth16Proof((0, 1, 2, 3), pk, a, b, c, conv(wa_vals), conv(wb_vals), cgroth16.fr_mont(r + 1), cgroth16.fr_mont(s),
                                       ec.SCALARS_MONTGOMERY if mont else ec.SCALARS_REGULAR)
        assert not np.array_equal(got2[0], got[0]) and not np.array_equal(got2[2], got[2]) and np.array_equal(got2[1], got[1])
        # errors: pk.G1.Z must have cardinality points; slots of the wrong kind
        with pytest.raises(ec.GkrB200EcError):
            ctx.ComputeGroth16Proof((0, 1, 0 if n_a < n else 3, 3), pk, a, b, c, conv(wa_vals), conv(wb_vals), cgroth16.fr_mont(r), cgroth16.fr_mont(s))
        with pytest.raises(ec.GkrB200EcError):
            ctx.ComputeGroth16Proof((0, 1, 2, 1), pk, a, b, c, conv(wa_vals), conv(wb_vals), cgroth16.fr_mont(r), cgroth16.fr_mont(s))
