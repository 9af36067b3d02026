"""gkr/ of the reference (prover side)."""
from ._lib import check, lib
from .context import _p, fr_array, fr_empty

PROOF_MONTGOMERY, PROOF_REGULAR = 0, 1


class Proof:
    """gkr.Proof{SumcheckProofs, Claims, QPrimes} (gkr/prover.go:14-18), decoded from the flat vector."""

    def __init__(self, c, bn, vec):
        self.bn, self.vec = bn, vec
        cur = 0
        self.SumcheckProofs, self.Claims, self.QPrimes = [], [], []
        for lay in c.layers:
            nco = 0 if lay.gate_kind is None else (3 if lay.gate_kind == 0 else 9)
            self.SumcheckProofs.append(vec[cur:cur + bn * nco].reshape(bn, nco, 4) if nco else None)
            cur += bn * nco
        for lay in c.layers:
            self.Claims.append(vec[cur:cur + len(lay.Out)])
            cur += len(lay.Out)
        for l, lay in enumerate(c.layers):
            nq = 1 if l == len(c.layers) - 1 else len(lay.Out)
            self.QPrimes.append(vec[cur:cur + nq * bn].reshape(nq, bn, 4))
            cur += nq * bn
        assert cur == vec.shape[0]

    def to_vec(self):
        """GkrProofToVec order (prover/gadget/hints.go:236-271)"""
        return self.vec


def Prove(c, a, qPrime, regular=False):
    """gkr.Prove(c, a, qPrime) (gkr/prover.go:21-47).  `a` must be the assignment currently held by c.ctx."""
    bn = a.bn
    q = fr_array(qPrime).reshape(-1, 4) if bn else None
    if bn and q.shape[0] != bn:
        raise ValueError("inconsistent sizes : qPrime has %d entries but bn is %d" % (q.shape[0], bn))
    vec = fr_empty(int(lib().gkrb200_proof_vec_len(bn)))
    check(lib().gkrb200_gkr_prove_mimc(c.ctx.handle, _p(q), bn, _p(vec), PROOF_REGULAR if regular else PROOF_MONTGOMERY))
    return Proof(c, bn, vec)


def Verify(c, a, proof, qPrime, regular=False, inputs=None, outputs=None):
    """gkr.Verify(c, proof, inputs, outputs, qPrime) (gkr/verifier.go:15-59).  Without inputs/outputs: against the assignment held by
    c.ctx (inputs = layers 0 and 1, outputs = layer 93, evaluated on the device).  With inputs=[key, msg] and outputs given (host
    tables, the reference's own arguments): evaluated from the caller's bytes, independent of the prover's assignment.
    Returns None when the proof is accepted, raises GkrB200Error otherwise (the reference returns an error value)."""
    bn = a.bn if a is not None else (len(qPrime) if qPrime is not None else 0)
    q = fr_array(qPrime).reshape(-1, 4) if bn else None
    vec = fr_array(proof.to_vec() if isinstance(proof, Proof) else proof)
    fl = PROOF_REGULAR if regular else PROOF_MONTGOMERY
    if inputs is None and outputs is None:
        check(lib().gkrb200_gkr_verify_mimc(c.ctx.handle, _p(vec), bn, _p(q), fl))
        return
    k, m, o = (fr_array(x).reshape(-1, 4) for x in (inputs[0], inputs[1], outputs))
    if not (k.shape[0] == m.shape[0] == o.shape[0] == 1 << bn):
        raise ValueError("inconsistent sizes : inputs/outputs must have 2^bn entries")
    check(lib().gkrb200_gkr_verify_mimc_io(c.ctx.handle, _p(vec), bn, _p(q), fl, _p(k), _p(m), _p(o)))
