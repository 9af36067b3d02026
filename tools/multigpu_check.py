"""Multi-GPU parity check, launched with torchrun (one rank per GPU): the sharded prover must produce the
byte-identical proof vector on every rank, equal to the oracle's, for every batch size incl. bn <= log2(world)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gkr-mimc_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import torch
import torch.distributed as dist
import gkrb200, coracle

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
max_bn = int(sys.argv[1]) if len(sys.argv) > 1 else 12
ctx = gkrb200.Context(local, max(max_bn, 4))
uid = [gkrb200.Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
leader = (world - 1) if len(sys.argv) > 2 and sys.argv[2] == "last" else 0
ctx.comm_init(rank, world, uid[0], leader=leader)
c = gkrb200.MimcCircuit(ctx)
ok = True
import time
# every size with ONE transcript per proof on the leader rank (default: window exchange, challenges consumed by device-side waits),
# then the lockstep modes: every rank runs the transcript, round sums through the window / through NCCL all-gather (A/B)
MODES = {0: "leader+window", 1: "lockstep+window", 2: "lockstep+nccl"}
sizes = sorted(set(list(range(0, 9)) + [max_bn]))
for mode, bn in [(0, b) for b in sizes] + [(1, b) for b in (2, 3, 5, 8, max_bn)] + [(2, b) for b in (2, 3, 5, 8, max_bn)]:
    if world > 1:
        ctx.set_option(gkrb200.Context.OPT_EXCHANGE, 1 if mode == 2 else 0)
        ctx.set_option(gkrb200.Context.OPT_TRANSCRIPT, 0 if mode == 0 else 1)
    n = 1 << bn
    rng = np.random.default_rng(100 + bn)
    key = gkrb200.common.RandomFrArray(n); msg = key[::-1].copy(); q = gkrb200.common.RandomFrArray(bn + 1)[1:]
    a = c.Assign(key, msg, want_outputs=True)
    t0 = time.perf_counter()
    vec = gkrb200.gkr.Prove(c, a, q).to_vec()
    prove_ms = (time.perf_counter() - t0) * 1e3
    out93, evec = coracle.assign_and_prove_mimc(key, msg, q)
    sharded = world > 1 and n > world
    exp_out = out93[rank::world] if sharded else out93
    good = np.array_equal(vec, evec) and np.array_equal(a.outputs, exp_out)
    if good and rank == 0:
        good = coracle.gkr_verify_mimc(vec, key, msg, out93, q) == 0
    if good:  # device-backed verifier and MLE evaluation of sharded layers (collective calls)
        try:
            gkrb200.gkr.Verify(c, a, vec, q)
            bad = vec.copy(); bad[3, 0] ^= np.uint64(1)
            try:
                gkrb200.gkr.Verify(c, a, bad, q); good = False
            except gkrb200.GkrB200Error:
                pass
            pt = gkrb200.common.RandomFrArray(bn + 3)[3:]
            full = coracle.mimc_assign(key, msg)
            for layer in (0, 50, 93):
                good = good and np.array_equal(a.Evaluate(layer, pt), coracle.evaluate(full[layer], pt))
        except gkrb200.GkrB200Error as e:
            print("rank %d: verifier rejected: %s" % (rank, e), flush=True); good = False
    t = torch.tensor([1 if good else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("bn=%2d world=%d sharded=%s mode=%s leader=%d prove=%.1f ms : %s" % (bn, world, sharded, MODES[mode], leader, prove_ms,
                                                                        "OK (bit-exact on all ranks)" if t.item() else "MISMATCH"), flush=True)
    ok = ok and bool(t.item())
ctx.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
