"""circuit/ + examples/ of the reference: the MiMC circuit and its device-resident assignment."""
from ._lib import check, lib
from .context import _p, fr_array, fr_empty

N_LAYERS = 94


class Layer:
    """circuit/circuit.go:14-23"""

    def __init__(self, In, gate_kind=None):
        self.In, self.Out, self.gate_kind = list(In), [], gate_kind


class Assignment:
    """circuit.Assignment (circuit/assignment.go:9): a[layer] reads the layer back from the device."""

    def __init__(self, ctx, n_local, bn):
        self.ctx, self.n_local, self.bn = ctx, n_local, bn

    def __len__(self):
        return N_LAYERS

    def __getitem__(self, layer):
        if layer < 0:
            layer += N_LAYERS
        out = fr_empty(self.n_local)
        check(lib().gkrb200_assign_layer_to_host(self.ctx.handle, layer, _p(out), self.n_local))
        return out

    def Evaluate(self, layer, coords):
        """a[layer].Evaluate(coords) (poly/multilin.go:59-66) without moving the layer off the device"""
        q = fr_array(coords).reshape(-1, 4) if self.bn else None
        out = fr_empty(1)
        check(lib().gkrb200_assign_layer_evaluate(self.ctx.handle, layer, _p(q), self.bn, _p(out)))
        return out[0]


class MimcCircuit:
    """examples.MimcCircuit() (examples/mimc.go:10-37) bound to a device context."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.layers = [Layer([]), Layer([]), Layer([0], 0)]
        for i in range(91):
            self.layers.append(Layer([2, 1 if i == 0 else i + 2], 1))
        for l, lay in enumerate(self.layers):  # circuit/circuit.go:28-44 BuildCircuit
            for pos in lay.In:
                self.layers[pos].Out.append(l)

    def __len__(self):
        return N_LAYERS

    def __getitem__(self, l):
        return self.layers[l]

    def check(self, arks=None):
        """gkrb200_check_mimc_circuit on this description (what the Go shim does before Assign / gkr.Prove): raises GkrB200Error
        unless it is examples.MimcCircuit().  arks: (94, 4) round constants per layer (default: hash.Arks in place)."""
        import ctypes
        import numpy as np
        from .common import Ark
        n_in = np.array([len(lay.In) for lay in self.layers], dtype=np.int32)
        flat = np.array([p for lay in self.layers for p in lay.In] or [0], dtype=np.int32)
        kinds = np.array([-1 if lay.gate_kind is None else lay.gate_kind for lay in self.layers], dtype=np.int32)
        if arks is None:
            arks = fr_empty(len(self.layers))
            for l in range(3, min(len(self.layers), N_LAYERS)):
                arks[l] = Ark(l - 3)
        arks = fr_array(arks)
        check(lib().gkrb200_check_mimc_circuit(len(self.layers), _p(n_in), _p(flat), _p(kinds), _p(arks)))

    def IsInputLayer(self, layer):
        return len(self.layers[layer].In) == 0

    def InputArity(self):
        return 2

    def _shape(self, n):
        world = self.ctx.world
        bn = n.bit_length() - 1
        sharded = world > 1 and (1 << bn) > world
        return bn, (n // world if sharded else n)

    def Assign(self, key, msg, want_outputs=False, out=None):
        """Circuit.Assign(inps...) (circuit/assignment.go:12-32).  key -> layer 0, msg -> layer 1.  want_outputs: a.outputs = a[93]
        (this rank's shard when multi-GPU), copied back by the same call; `out` = caller's buffer for it (e.g. pinned memory)."""
        k = fr_array(key).reshape(-1, 4)
        m = fr_array(msg).reshape(-1, 4)
        if k.shape != m.shape:
            raise ValueError("inputs must have the same length")
        n = k.shape[0]
        bn, n_local = self._shape(n)
        out93 = None
        if out is not None:
            out93 = fr_array(out).reshape(-1, 4)[:n_local]
            want_outputs = True
        elif want_outputs:
            out93 = fr_empty(n_local)
        check(lib().gkrb200_mimc_assign(self.ctx.handle, _p(k), _p(m), n, _p(out93)))
        a = Assignment(self.ctx, n_local, bn)
        if want_outputs:
            a.outputs = out93
        return a

    IO_INPUT_REGULAR, IO_OUTPUT_REGULAR, IO_OUTPUT_HASH = 1, 2, 4

    def AssignEx(self, key, msg, flags):
        """The hint's view of Assign (prover/gadget/hints.go:135-145,197-233): inputs optionally in regular form (SetBigInt on the
        device), and `outputs` = layer 93 or the gadget's hash a[93] + 2*key + msg, optionally in regular form."""
        k = fr_array(key).reshape(-1, 4)
        m = fr_array(msg).reshape(-1, 4)
        if k.shape != m.shape:
            raise ValueError("inputs must have the same length")
        n = k.shape[0]
        bn, n_local = self._shape(n)
        out = fr_empty(n_local)
        check(lib().gkrb200_mimc_assign_ex(self.ctx.handle, _p(k), _p(m), n, _p(out), flags))
        a = Assignment(self.ctx, n_local, bn)
        a.outputs = out
        return a

    def AssignDevice(self, d_key_ptr, d_msg_ptr, n):
        """Same with inputs already on the device (raw device pointers, Go layout); asynchronous."""
        bn, n_local = self._shape(n)
        check(lib().gkrb200_mimc_assign_device(self.ctx.handle, d_key_ptr, d_msg_ptr, n))
        return Assignment(self.ctx, n_local, bn)
