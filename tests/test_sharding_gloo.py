"""world_size 2/4/8 CPU tests (gloo) of the multi-GPU decomposition the library implements with NCCL (SURVEY.md
section 5/8e, gkrb200.cu Ctx::sumcheck): rank g owns table entries {i : i mod G == g}; its eq shard is
s_g * eq(q[0:bn-log2 G], .); per round the ranks exchange only their partial round evaluations; after
bn-log2 G rounds the G residual entries are gathered and the last rounds are finished identically everywhere.
The big-int restatement (oracle/pyref.py) plays the part of the kernels; the exchange runs over torch.distributed."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _to_t(vals):
    """field elements as 5 x 62-bit limbs in an int64 tensor (gloo has no 256-bit dtype)"""
    return torch.tensor([[(v >> (62 * j)) & ((1 << 62) - 1) for j in range(5)] for v in vals], dtype=torch.int64)


def _from_t(t):
    return [sum(int(x) << (62 * j) for j, x in enumerate(row)) for row in t.tolist()]


def _worker(rank, world, port, bn, kind, q, out):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyref as P
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        logw = world.bit_length() - 1
        bnl = bn - logw
        n = 1 << bn
        gate = P.Gate(kind, 145646)
        L = [(i * i + 3) % P.Q for i in range(n)]
        Rr = [(7 * i + 11) % P.Q for i in range(n)]
        tabs = [L, Rr] if kind == "cipher" else [L]
        # strided shards (gkrb200_mimc_assign keeps entries i = j*G + rank)
        X = [t[rank::world] for t in tabs]
        s = 1
        for b in range(logw):  # address bit b of the rank pairs with q[bn-1-b]
            qb = q[bn - 1 - b]
            s = s * (qb if (rank >> b) & 1 else (1 - qb)) % P.Q
        eq = P.folded_eq_table(q[:bnl], s)
        full_eq = P.folded_eq_table(q)
        assert eq == full_eq[rank::world]
        proof, chal = [], []

        def allgather(vals):
            mine = _to_t(vals)
            buf = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(buf, mine)
            return [_from_t(b) for b in buf]

        for _ in range(bnl):
            part = P.partial_evals(eq, X, gate)
            ev = [sum(col) % P.Q for col in zip(*allgather(part))]  # modular sum of the per-rank partials
            co = P.interpolate_on_range(ev)
            r = P.get_challenge(co)  # every rank derives the same challenge: no broadcast needed
            eq = P.fold(eq, r)
            X = [P.fold(x, r) for x in X]
            proof.append(co)
            chal.append(r)
        resid = allgather([eq[0]] + [x[0] for x in X])  # entry g of the residual tables is rank g's value
        eq = [resid[g][0] for g in range(world)]
        X = [[resid[g][1 + k] for g in range(world)] for k in range(len(X))]
        for _ in range(logw):
            co = P.interpolate_on_range(P.partial_evals(eq, X, gate))
            r = P.get_challenge(co)
            eq = P.fold(eq, r)
            X = [P.fold(x, r) for x in X]
            proof.append(co)
            chal.append(r)
        fin = [eq[0]] + [x[0] for x in X]
        ref = P.sumcheck_prove(tabs, [q], [], gate)
        assert (proof, chal, fin) == ref, "sharded sumcheck differs from the single-device one"
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,bn,world", [("cipher", 5, 2), ("identity", 4, 2), ("cipher", 1, 2), ("cipher", 5, 4), ("identity", 4, 8), ("cipher", 3, 8)])
def test_sharded_sumcheck_equals_single(kind, bn, world):
    """2, 4 and 8 ranks (the box sizes bench.py runs), incl. bn == log2(world): every rank holds ONE entry per table"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyref as P
    q = P.random_fr_array(bn + 2)[2:]
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, bn, kind, q, out), nprocs=world, join=True)
    assert dict(out) == {g: 1 for g in range(world)}
