// ComputeGroth16Proof (prover/gadget/prove.go:100-306) as a sequence of the library's device operations.
//
// The reference runs computeH, three G1 multi-exponentiations, one G2 multi-exponentiation and a handful of single-point operations
// in goroutines; here the same values are produced by one sequence over an operations backend `Ops`:
//   msm_g1 / msm_g2       G1Affine / G2Affine.MultiExp against a resident base slot
//   msm_g1_h              the same against the h that computeH left on the device (krs2, prove.go:221)
//   smul_g1 / smul_g2     ScalarMultiplication of one point (BatchScalarMultiplicationG1 for the three deltas, prove.go:176;
//                         s * ar, r * bs1, prove.go:246,254; s * pk.G2.Delta, prove.go:284) -- a one-point multi-exponentiation
//   add_g1 / add_g2       AddMixed / AddAssign; every intermediate is affine here (the reference keeps Jacobian accumulators;
//                         the group elements, and therefore the affine outputs, are the same)
// The CUDA backend is in ec.cu; tests/emu/ec_emu.cpp provides a host backend over the emulated kernels, so this sequencing (the
// only logic in it is kr = -(r s) and the order of the additions) is checked on the CPU against the oracle's composition.
// r and s are sampled by the CALLER (prove.go:154-161 uses fr.SetRandom: that stays in Go), which keeps this function a
// deterministic map from its inputs to (Ar, Bs, Krs).
#pragma once
#include "field.cuh"

namespace ec {

struct Groth16Out {
    uint64_t ar[8], bs[16], krs[8];
};

// r_mont, s_mont: fr.Element images (Montgomery).  Returns 0 or the first non-zero status of an operation.
template <class Ops>
int groth16_compose(Ops& o, const uint64_t* r_mont, const uint64_t* s_mont, Groth16Out& out) {
    // _kr = -(r * s); then all three in regular form (prove.go:162-168)
    const Big8 rm = big_load(r_mont), sm = big_load(s_mont);
    const Big8 kr = f_neg<FrMod>(f_mul<FrMod>(rm, sm));
    uint64_t r[4], s[4], krr[4];
    big_store(r, f_from_mont<FrMod>(rm));
    big_store(s, f_from_mont<FrMod>(sm));
    big_store(krr, f_from_mont<FrMod>(kr));
    int rc;
    // deltas := BatchScalarMultiplicationG1(&pk.G1.Delta, {r, s, kr})   (prove.go:176)
    uint64_t d_r[8], d_s[8], d_kr[8];
    if ((rc = o.smul_g1(o.delta(), r, d_r))) return rc;
    if ((rc = o.smul_g1(o.delta(), s, d_s))) return rc;
    if ((rc = o.smul_g1(o.delta(), krr, d_kr))) return rc;
    // h = computeH(a, b, c)   (prove.go:126); stays on the device
    if ((rc = o.compute_h())) return rc;
    // bs1 = MultiExp(pk.G1.B, wireValuesB) + Beta + deltas[1]   (prove.go:186-196)
    uint64_t bs1[8], t[8];
    if ((rc = o.msm_g1(1, t))) return rc;
    if ((rc = o.add_g1(t, o.beta(), bs1))) return rc;
    if ((rc = o.add_g1(bs1, d_s, bs1))) return rc;
    // ar = MultiExp(pk.G1.A, wireValuesA) + Alpha + deltas[0]   (prove.go:199-210)
    uint64_t ar[8];
    if ((rc = o.msm_g1(0, t))) return rc;
    if ((rc = o.add_g1(t, o.alpha(), ar))) return rc;
    if ((rc = o.add_g1(ar, d_r, ar))) return rc;
    // krs = deltas[2] + MultiExp(pk.G1.Z, h) + s * ar + r * bs1   (prove.go:212-262; the pk.G1.K term is handled by the caller, :224-231)
    uint64_t krs[8], p1[8];
    if ((rc = o.msm_g1_h(t))) return rc;
    if ((rc = o.add_g1(d_kr, t, krs))) return rc;
    if ((rc = o.smul_g1(ar, s, p1))) return rc;
    if ((rc = o.add_g1(krs, p1, krs))) return rc;
    if ((rc = o.smul_g1(bs1, r, p1))) return rc;
    if ((rc = o.add_g1(krs, p1, krs))) return rc;
    // Bs = MultiExp(pk.G2.B, wireValuesB) + s * pk.G2.Delta + pk.G2.Beta   (prove.go:265-292)
    uint64_t bs[16], t2[16];
    if ((rc = o.msm_g2(bs))) return rc;
    if ((rc = o.smul_g2(o.delta2(), s, t2))) return rc;
    if ((rc = o.add_g2(bs, t2, bs))) return rc;
    if ((rc = o.add_g2(bs, o.beta2(), bs))) return rc;
    for (int i = 0; i < 8; i++) out.ar[i] = ar[i], out.krs[i] = krs[i];
    for (int i = 0; i < 16; i++) out.bs[i] = bs[i];
    return 0;
}

}  // namespace ec
