"""gkrb200 -- host-side mirror of the gkr-mimc prover interface over the B200 C-ABI library.

The names follow the reference's Go packages so parity tests read like the reference's own tests:

    reference (Go)                                  here
    ----------------------------------------------  -----------------------------------------
    examples.MimcCircuit()                          gkrb200.MimcCircuit(ctx)
    circuit.Circuit.Assign(key, msg) -> Assignment  MimcCircuit.Assign(key, msg) -> Assignment
    gkr.Prove(c, a, qPrime) -> gkr.Proof            gkrb200.gkr.Prove(c, a, qPrime) -> Proof
    sumcheck.Prove(X, qPrimes, claims, gate)        gkrb200.sumcheck.Prove(ctx, X, qPrimes, claims, gate)
    poly.FoldedEqTable / MultiLin.Fold              gkrb200.poly.FoldedEqTable / Fold
    common.GetChallenge                             gkrb200.common.GetChallenge
    gates.IdentityGate{} / gates.NewCipherGate(ark) gkrb200.gates.IdentityGate() / CipherGate(ark)
    GkrProofToVec (prover/gadget/hints.go:236)      Proof.to_vec()

Field elements are numpy uint64 arrays (..., 4): Go's []fr.Element image (Montgomery, LE limbs).
Everything that touches a table runs on the GPU through libgkrb200.so; there is no CPU fallback.
"""
from ._lib import GkrB200Error, Stats, build, lib  # noqa: F401
from .context import Context  # noqa: F401
from . import common, gates, gkr, poly, sumcheck  # noqa: F401
from .circuit import Assignment, MimcCircuit  # noqa: F401
