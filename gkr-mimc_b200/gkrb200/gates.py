"""circuit/gates of the reference.  Only the two gates of the MiMC circuit can cross the C ABI
(the Go shim type-switches on them)."""
import numpy as np

GATE_IDENTITY, GATE_CIPHER = 0, 1


class IdentityGate:
    """circuit/gates/copy.go:9-32"""
    kind = GATE_IDENTITY
    ark = None

    def ID(self):
        return "CopyGate"

    def Degree(self):
        return 1


class CipherGate:
    """circuit/gates/cipher.go:11-70: (vL + vR + Ark)^7"""
    kind = GATE_CIPHER

    def __init__(self, ark):
        self.ark = np.ascontiguousarray(ark, dtype=np.uint64).reshape(4)

    def ID(self):
        return "CipherGate"

    def Degree(self):
        return 7


def NewCipherGate(ark):
    return CipherGate(ark)
