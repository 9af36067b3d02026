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


def _free_port():
    """a TCP port the kernel just handed out as free (avoids collisions between consecutive rendezvous on a fixed port)"""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


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
    port = _free_port()
    mp.spawn(_worker, args=(world, port, bn, kind, q, out), nprocs=world, join=True)
    assert dict(out) == {g: 1 for g in range(world)}


def _upload_worker(rank, world, port, n, out):
    """stage_inputs in gkrb200.cu: each rank holds only its CONTIGUOUS 1/world slice of the table, de-interleaves it by owner
    (k_destripe: send[d][j] = slice[j*world + d]) and the ranks swap the blocks (NCCL send/recv there, gloo here)."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        table = torch.arange(n, dtype=torch.int64) * 1000003 + 17
        nl = n // world
        blk = nl // world
        mine = table[rank * nl:(rank + 1) * nl]
        send = torch.stack([mine[d::world] for d in range(world)])  # send[d][j] = slice[j*world + d]
        assert send.shape == (world, blk)
        recv = torch.zeros((world, blk), dtype=torch.int64)
        reqs = []
        for peer in range(world):
            if peer == rank:
                recv[peer] = send[peer]
                continue
            reqs.append(dist.isend(send[peer].contiguous(), dst=peer))
            reqs.append(dist.irecv(recv[peer], src=peer))
        for r in reqs:
            r.wait()
        local = recv.reshape(-1)  # block from rank g lands at local[g*blk ...]
        assert torch.equal(local, table[rank::world]), "re-strided shard differs from table[rank::world]"
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 16), (4, 64), (8, 64), (8, 512)])
def test_sharded_upload_all_to_all(world, n):
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_upload_worker, args=(world, port, n, out), nprocs=world, join=True)
    assert dict(out) == {g: 1 for g in range(world)}


def _leader_worker(rank, world, port, bn, leader, q, out):
    """One transcript per proof (gkrb200.cu lead_mode): every rank evaluates its shard, ONLY the leader interpolates and hashes,
    the challenge travels back to all ranks (the exchange window + device-side wait on the GPUs, a broadcast here), followers
    never call get_challenge; at the end the leader's layer header gives every rank the same proof."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyref as P
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        logw = world.bit_length() - 1
        bnl = bn - logw
        n = 1 << bn
        gate = P.Gate("cipher", 145646)
        tabs = [[(i * i + 3) % P.Q for i in range(n)], [(7 * i + 11) % P.Q for i in range(n)]]
        X = [t[rank::world] for t in tabs]
        s = 1
        for b in range(logw):
            qb = q[bn - 1 - b]
            s = s * (qb if (rank >> b) & 1 else (1 - qb)) % P.Q
        eq = P.folded_eq_table(q[:bnl], s)
        hashes = 0

        def to_leader(vals):
            mine = _to_t(vals)
            buf = [torch.zeros_like(mine) for _ in range(world)] if rank == leader else None
            dist.gather(mine, buf, dst=leader)
            return [_from_t(b) for b in buf] if rank == leader else None

        def from_leader(vals, k):
            t = _to_t(vals) if rank == leader else torch.zeros((k, 5), dtype=torch.int64)
            dist.broadcast(t, src=leader)
            return _from_t(t)

        proof, chal = [], []
        for _ in range(bnl):
            parts = to_leader(P.partial_evals(eq, X, gate))
            r = None
            if rank == leader:
                co = P.interpolate_on_range([sum(col) % P.Q for col in zip(*parts)])
                r = P.get_challenge(co)
                hashes += 1
                proof.append(co)
                chal.append(r)
            r = from_leader([r] if rank == leader else None, 1)[0]
            eq = P.fold(eq, r)
            X = [P.fold(x, r) for x in X]
        resid = to_leader([eq[0]] + [x[0] for x in X])
        fin = None
        if rank == leader:  # host tail on the leader only
            eq = [resid[g][0] for g in range(world)]
            X = [[resid[g][1 + k] for g in range(world)] for k in range(2)]
            for _ in range(logw):
                co = P.interpolate_on_range(P.partial_evals(eq, X, gate))
                r = P.get_challenge(co)
                hashes += 1
                eq = P.fold(eq, r)
                X = [P.fold(x, r) for x in X]
                proof.append(co)
                chal.append(r)
            fin = [eq[0]] + [x[0] for x in X]
        # the layer header: round polynomials, challenges, final claims -> every rank
        flat = from_leader([c for co in proof for c in co] + chal + fin if rank == leader else None, bn * 9 + bn + 3)
        got = ([flat[9 * k:9 * k + 9] for k in range(bn)], flat[9 * bn:10 * bn], flat[10 * bn:])
        assert got == P.sumcheck_prove(tabs, [q], [], gate), "leader-transcript sumcheck differs from the single-device one"
        assert hashes == (bn if rank == leader else 0)
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bn,world,leader", [(5, 2, 0), (5, 2, 1), (4, 4, 3), (3, 8, 5)])
def test_leader_transcript_equals_single(bn, world, leader):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyref as P
    q = P.random_fr_array(bn + 2)[2:]
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_leader_worker, args=(world, port, bn, leader, q, out), nprocs=world, join=True)
    assert dict(out) == {g: 1 for g in range(world)}
