#!/usr/bin/env python3
"""Regenerates the golden fixtures from the pure-Python big-int restatement (oracle/pyref.py).

The reference is Go and cannot run in this image (no Go toolchain), and it holds no golden proof bytes of
its own; these fixtures therefore come from oracle/pyref.py, which is pinned on the reference's own
known-answer tests (hash/hash_test.go:21-27 etc., see tests/test_oracle.py).  They agree with the digests
SURVEY.md appendix B recorded from an earlier, independently written restatement.

  gkr_proof_digests.json : sha256 over GkrProofToVec (regular form, 32-byte big-endian words) for
                           gkr/gkr_test.go:23-25 inputs, bn = 0..5
  sumcheck_cipher.json   : InitializeCipherGateInstance(bn) (sumcheck/testing.go:11-26), bn = 1..4:
                           claim, sha256 of all round coefficients, first challenge
  kat.json               : MimcHash([12]) (reference golden), MimcHash([0]), RandomFrArray(4), a[93][0]
"""
import hashlib
import json
import os
import sys

here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(here, "..", "..", "oracle"))
import pyref as P  # noqa: E402


def sha(vals):
    return hashlib.sha256(b"".join(v.to_bytes(32, "big") for v in vals)).hexdigest()


def main():
    c = P.mimc_circuit()
    gkr = {}
    a93 = None
    for bn in range(6):
        blk = P.random_fr_array(1 << bn)
        qp = P.random_fr_array(bn)
        a = P.assign(c, blk, blk)
        a93 = a[93][0]
        pr = P.gkr_prove(c, a, qp)
        P.gkr_verify(c, pr, [blk, blk], a[93], qp)
        gkr[str(bn)] = sha(P.gkr_proof_to_vec(pr))
    json.dump(gkr, open(os.path.join(here, "gkr_proof_digests.json"), "w"), indent=1)
    sc = {}
    for bn in range(1, 5):
        X, cl, q, g = P.init_cipher_gate_instance(bn)
        proof, ch, _ = P.sumcheck_prove(X, q, cl, g)
        sc[str(bn)] = {"claim": str(cl[0]), "sha256_coeffs": sha([x for rnd in proof for x in rnd]), "first_challenge": str(ch[0])}
    json.dump(sc, open(os.path.join(here, "sumcheck_cipher.json"), "w"), indent=1)
    kat = {"mimc_hash_12": str(P.mimc_hash([12])), "mimc_hash_0": str(P.mimc_hash([0])),
           "random_fr_array_4": [str(x) for x in P.random_fr_array(4)], "a93_0": str(a93)}
    json.dump(kat, open(os.path.join(here, "kat.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
