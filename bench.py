#!/usr/bin/env python3
"""bench.py -- proven MiMC hashes/sec (bit-exact GKR proof) on N B200s.

One "step" = one full pass of the hot path over one batch: Circuit.Assign (circuit/assignment.go:12) +
gkr.Prove (gkr/prover.go:21) of a 2^bn-hash batch through the C ABI of libgkrb200.so.

  python bench.py --gpus 1 --steps 5 --warmup 3                 # BASELINE config 5 (2^22 hashes), the headline
  python bench.py --config 3                                    # configs 1/3/4/5 = 2^10 / 2^16 / 2^20 / 2^22 full proofs
  python bench.py --config 2                                    # standalone sumcheck of eq*gate over 2^20-entry tables
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
  python bench.py --impl reference ...      # the CPU port of the reference's Go prover on the host cores

`value`  : whole-job hashes/s with the inputs already resident in HBM when the timed region starts.
`e2e`    : same metric through the host-buffer API (gkrb200_mimc_assign + gkrb200_gkr_prove_mimc): H2D of the inputs, D2H of
           the hash outputs a[93] and of the proof inside the timed region, host buffers pinned.
`roofline`: the dominant kernel (k_round_cf: fold + factored round evaluation) timed live with CUDA events around every
           launch of one extra step: algorithmic field multiplications / that time against the measured integer-multiply
           peak (the bound of this kernel), plus `roofline_hbm` (algorithmic bytes / that time against the measured HBM peak).
`parity` : the GPU arm proves the batch the cpu_baseline leg proves (same seeds) and compares sha256 of outputs and proof.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "gkr-mimc_b200"))
# Every proof in flight owns a stream on which a whole layer of launches is pre-enqueued, each waiting on the previous one.  With the
# driver's default of 8 hardware work queues several streams share a queue, and a launch that is waiting for its predecessor holds
# back the independent launches of another proof queued behind it (false dependency).  32 queues = one per stream.  Must be set
# before CUDA initialises.  Single-GPU runs launch one round at a time and keep the driver's default.
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_JSON_OUT = sys.stdout
METRIC = "proven MiMC hashes/sec (bit-exact GKR proof)"
UNIT = "hashes/s"
Q3 = 0x30644E72E131A029  # top limb of q: any element with top limb < Q3 is canonical
CONFIG_BN = {1: 10, 3: 16, 4: 20, 5: 22}  # BASELINE.json configs that are full GKR proofs
CONFIG_NAME = {
    1: "config 1: examples/mimc_test -- GKR-MiMC proof for a 2^10-hash batch",
    2: "config 2: standalone sumcheck (sumcheck/prover_test.go:96-125 style) of eq*gate over 2^20-entry L/R tables",
    3: "config 3: full MiMC GKR proof for a 2^16-hash batch",
    4: "config 4: full MiMC GKR proof for a 2^20-hash batch",
    5: "config 5: full MiMC GKR proof for a 2^22-hash batch, all layer assignments resident in HBM",
}


def synth_inputs(n, seed):
    """Deterministic pseudo-random canonical field elements (Go layout). Seeds are recorded in the JSON line."""
    import numpy as np
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    a[:, :3] |= rng.integers(0, 2, size=(n, 3), dtype=np.uint64) << np.uint64(63)
    a[:, 3] %= np.uint64(Q3)
    return a


def sha(a):
    import numpy as np
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md 'clocks line')."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU port (oracle/) legs
def _coracle(threads):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import coracle
    coracle.build()
    coracle.set_threads(threads)
    return coracle


def cpu_port_prove(bn, threads, seed, inputs=None):
    """Times the CPU port of the reference prover (oracle/, test infrastructure) on a 2^bn batch.  Input synthesis and the
    build check are OUTSIDE the timed region.  Returns (hashes/s, seconds, out93, proof_vec)."""
    coracle = _coracle(threads)
    n = 1 << bn
    key, msg, qp = inputs if inputs is not None else (synth_inputs(n, seed), synth_inputs(n, seed + 1), synth_inputs(bn, seed + 2))
    t0 = time.perf_counter()
    out93, vec = coracle.assign_and_prove_mimc(key, msg, qp)
    dt = time.perf_counter() - t0
    return n / dt, dt, out93, vec


def cpu_port_sumcheck(bn, threads, seed):
    """standalone cipher-gate sumcheck (config 2) on the CPU port: returns (entries/s, seconds)"""
    import numpy as np
    coracle = _coracle(threads)
    n = 1 << bn
    L, R, q = synth_inputs(n, seed), synth_inputs(n, seed + 1), synth_inputs(bn, seed + 2).reshape(1, bn, 4)
    ark = synth_inputs(1, seed + 3)[0]
    claim = np.zeros((1, 4), dtype=np.uint64)  # the prover does not read a single claim (sumcheck/prover.go:121-141)
    t0 = time.perf_counter()
    coracle.sumcheck_prove([L, R], q, claim, coracle.GATE_CIPHER, ark)
    dt = time.perf_counter() - t0
    return n / dt, dt


def fr_mul_ns(threads=1):
    coracle = _coracle(threads)
    thr, lat = coracle.bench_fr_mul(2_000_000)
    return {"ns_per_fr_mul_single_thread": thr, "ns_per_dependent_fr_mul": lat, "multiplier": coracle.fr_mul_kind()}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The Go toolchain does not exist in
    this image, so this is the C port of the Go prover (oracle/gkr_oracle.c: same task decomposition, one worker per core,
    MULX/ADX multiplier of the class gnark-crypto's assembly is).  Each step proves ONE 2^ref_bn batch, a bounded sample of the
    named 2^bn workload: ref_bn is the largest size <= min(bn, --ref-bn) for which steps x (measured time) fits --ref-budget-s;
    the line's `config` states the batch that was actually run."""
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    mul = fr_mul_ns()
    if args.config == 2:
        bn = min(args.ref_bn, args.bn)
        for _ in range(max(args.warmup, 1)):
            cpu_port_sumcheck(min(bn, 12), cores, args.seed)
        t_all, done = 0.0, 0
        for _ in range(args.steps):
            _, dt = cpu_port_sumcheck(bn, cores, args.seed)
            t_all += dt
            done += 1 << bn
        value, ms_step = done / t_all, t_all / args.steps * 1e3
        sample = "each step = one standalone cipher-gate sumcheck.Prove over 2^%d-entry tables, C port of the Go prover, %d threads" % (bn, cores)
        metric, unit = "standalone sumcheck (eq * cipher gate) table entries/sec", "entries/s"
    else:
        # pick the sample size from a measured probe so that the whole run stays inside the budget
        for _ in range(max(args.warmup, 1)):
            cpu_port_prove(min(args.bn, 12), cores, args.seed)
        probe_bn = min(args.bn, 16)
        rate, _, _, _ = cpu_port_prove(probe_bn, cores, args.seed)
        bn = min(args.bn, args.ref_bn)
        while bn > probe_bn and args.steps * ((1 << bn) / rate) > args.ref_budget_s:
            bn -= 1
        inputs = (synth_inputs(1 << bn, args.seed), synth_inputs(1 << bn, args.seed + 1), synth_inputs(bn, args.seed + 2))
        t_all, done = 0.0, 0
        for _ in range(args.steps):
            _, dt, _, _ = cpu_port_prove(bn, cores, args.seed, inputs)
            t_all += dt
            done += 1 << bn
        value, ms_step = done / t_all, t_all / args.steps * 1e3
        sample = "each step = full Assign+Prove of ONE 2^%d-hash batch%s, C port of the Go prover, %d threads" % (
            bn, "" if bn == args.bn else " (bounded sample of the named 2^%d workload; the port's hashes/s grows slowly with the batch)" % args.bn, cores)
        metric, unit = METRIC, UNIT
    cfg = workload_config(args, world)
    cfg.update({"bn": bn, "hashes_per_step": 1 << bn, "proof_elements": 1006 * bn + 183, "named_workload_bn": args.bn, "same_batch_as_b200_arm": bn == args.bn,
                "parallelism": "%d host threads (one worker per core, sumcheck/worker.go:14-26)" % cores,
                "workload": ("standalone cipher-gate sumcheck over 2^%d-entry tables" % bn) if args.config == 2 else
                "full MiMC GKR proof (Circuit.Assign + gkr.Prove, 94 layers) of a 2^%d-hash batch over BN254 Fr on the host CPU" % bn})
    cfg.pop("l2", None)
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64x4 (BN254 Fr, Montgomery)",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample, **mul},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    return 0


# dram__bytes_read.sum + dram__bytes_write.sum of the largest k_round_cf launches (ncu --set full, profiles/): round 0 (2^21 pairs, no
# fold) vs 268.4 MB algorithmic
NCU_TRAFFIC = {22: 272.7e6}
NCU_TRAFFIC_NOTE = "ncu capture of the round-0 launch of one layer (2^21 pairs: 268.4 MB algorithmic); `achieved` averages all launches of a proof"


def workload_config(args, world, P=1, replicas=False, exchange=None):
    xdesc = {"window": "round sums published by the round kernels into a host-shared mapped window, one transcript per proof on its leader rank, "
                       "challenges consumed by device-side waits (no collective on the round path)",
             "nccl": "NCCL all-gather of the round sums"}.get(exchange, "round sums exchanged every round")
    return {
        "workload": "full MiMC GKR proof (Circuit.Assign + gkr.Prove, 94 layers, transcript bit-exact) of a 2^%d-hash batch over BN254 Fr" % args.bn,
        "baseline_config": CONFIG_NAME.get(args.config),
        "bn": args.bn, "hashes_per_step": (1 << args.bn) * (world if replicas else 1), "proof_elements": 1006 * args.bn + 183,
        "parallelism": "single GPU, %d proofs in flight" % P if world == 1 else (
            "replicas: each of the %d GPUs proves its own 2^%d batches, no exchange, %d proofs in flight per GPU" % (world, args.bn, P) if replicas else
            "batch sharded on low address bits over %d GPUs, %s, %d proofs in flight" % (world, xdesc, P)),
        "l2": "inputs larger than L2: 93 layer tables of %d MiB each per proof" % ((32 << args.bn) >> 20) if args.bn >= 18 else
              "L2 flushed between timed steps (a 256 MiB buffer is overwritten): the working set of a 2^%d batch fits in L2" % args.bn,
        "seeds": [args.seed, args.seed + 1, args.seed + 2],
        **({"options": args.opt} if args.opt else {}),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=[1, 2, 3, 4, 5], help="BASELINE.json config (1-based); the metric is quoted on config 5 (2^22 hashes)")
    ap.add_argument("--bn", type=int, default=None, help="log2 of the batch (overrides the config's size)")
    ap.add_argument("--ref-bn", type=int, default=20, help="largest batch of one --impl reference step (bounded sample)")
    ap.add_argument("--ref-budget-s", type=float, default=300.0, help="--impl reference: the timed steps together must fit this many seconds")
    ap.add_argument("--cpu-bn", type=int, default=20, help="batch of the cpu_baseline sample and of the parity check")
    ap.add_argument("--seed", type=int, default=0x6B6B72)
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg and the parity check against it")
    ap.add_argument("--inflight", type=int, default=0, help="proofs in flight per GPU (0 = auto: up to 8, bounded by --steps and by the host cores)")
    ap.add_argument("--mode", default="sharded", choices=["sharded", "replicas"],
                    help="N > 1 GPUs: 'sharded' = every 2^bn batch is split over all N GPUs (one proof, round sums exchanged; the north-star "
                         "configuration, strong scaling); 'replicas' = every GPU proves its own 2^bn batches (no exchange, weak scaling)")
    ap.add_argument("--opt", action="append", default=[], metavar="ID=VALUE", help="gkrb200_set_option passthrough (tuning experiments), repeatable")
    args = ap.parse_args()
    if args.bn is None:
        args.bn = 20 if args.config == 2 else CONFIG_BN[args.config]
    args.cpu_bn = min(args.cpu_bn, args.bn)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner) go to stderr instead
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    import gkrb200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    env = dict(rank=rank, local_rank=local_rank, world=world, dist=dist, barrier=barrier, max_over_ranks=max_over_ranks, sum_over_ranks=sum_over_ranks)
    if args.config == 2:
        return run_sumcheck_config(args, env)

    n, bn = 1 << args.bn, args.bn
    # P proofs in flight: each has its own context (arena + stream) and its own host thread, so the serial host transcript of one
    # proof (MiMC challenges, ~110 ms per 2^22 proof) overlaps the device rounds of another.  Sharded mode: ONE host thread per
    # proof runs the transcript (its leader rank, rotating over the ranks); the other ranks' threads only enqueue kernels and
    # sleep, so the proofs in flight are bounded by the host cores of the box, not by cores / world.
    replicas = world > 1 and args.mode == "replicas"
    sharded = world > 1 and not replicas
    cores = os.cpu_count() or 1
    if args.inflight > 0:
        P = args.inflight
    elif replicas:
        P = max(1, min(8, args.steps, cores // world))
    elif sharded:
        # one spinning/hashing host thread per proof in flight on the whole box (its leader); the other ranks' threads sleep.
        # 8 in flight is the configuration validated on 2 and 8 x B200 (profiles/r2_bench_bn22_n{2,8}_leader*.json); a first attempt
        # with 24 in flight on 8 GPUs did not finish (profiles/r2_note_p24_n8.txt), so more is opt-in (--inflight).
        P = max(1, min(8, args.steps, cores - 2))
    else:
        P = max(1, min(8, args.steps, cores - 2))
    main_stream = torch.cuda.Stream()
    streams = [torch.cuda.Stream() for _ in range(P)]  # the library launches on these; events are recorded on main_stream after joining them
    torch.cuda.set_stream(main_stream)
    # more than 8 sharded contexts per GPU only fit when their arenas are sized for the shard from the start (gkrb200_init_shard)
    ctxs = [gkrb200.Context(device=local_rank, max_bn=bn, stream=st_.cuda_stream, world=world if (sharded and P > 8) else 1) for st_ in streams]
    ctx = ctxs[0]
    if sharded:
        for i, c in enumerate(ctxs):  # one communicator per pipeline, created in the same order on every rank
            uid = [gkrb200.Context.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            c.comm_init(rank, world, uid[0], leader=i % world)
    for kv in args.opt:  # after comm_init: GKRB200_OPT_EXCHANGE refers to the communicator
        oid, val = kv.split("=")
        for c in ctxs:
            c.set_option(int(oid), int(val))
    circuits = [gkrb200.MimcCircuit(c) for c in ctxs]

    # synthetic inputs: pinned host copies (e2e path) and device-resident copies (kernel path)
    key_np, msg_np, q_np = synth_inputs(n, args.seed), synth_inputs(n, args.seed + 1), synth_inputs(bn, args.seed + 2)
    key_h = torch.from_numpy(key_np.view(np.int64)).pin_memory()
    msg_h = torch.from_numpy(msg_np.view(np.int64)).pin_memory()
    key_d, msg_d = key_h.cuda(non_blocking=True), msg_h.cuda(non_blocking=True)
    key_hn, msg_hn = key_h.numpy().view(np.uint64), msg_h.numpy().view(np.uint64)
    n_local = n // world if (sharded and n > world) else n
    out_h = [torch.empty((n_local, 4), dtype=torch.int64).pin_memory() for _ in range(P)]  # a[93] lands here in the e2e path
    out_hn = [o.numpy().view(np.uint64) for o in out_h]
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if bn < 18 else None
    torch.cuda.synchronize()

    def step_resident(i=0):
        if flush_buf is not None:
            with torch.cuda.stream(streams[i]):
                flush_buf.zero_()
        a = circuits[i].AssignDevice(key_d.data_ptr(), msg_d.data_ptr(), n)
        return gkrb200.gkr.Prove(circuits[i], a, q_np)

    def step_e2e(i=0):
        if flush_buf is not None:
            with torch.cuda.stream(streams[i]):
                flush_buf.zero_()
        a = circuits[i].Assign(key_hn, msg_hn, out=out_hn[i])  # H2D of key/msg, D2H of the hash outputs a[93]
        return gkrb200.gkr.Prove(circuits[i], a, q_np)

    def timed(fn, steps, workers):
        """exactly `steps` proofs, spread over `workers` concurrent pipelines; device time between two events on main_stream"""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        counts = [steps // workers + (1 if i < steps % workers else 0) for i in range(workers)]
        outs, errs = [None] * workers, []

        def work(i):
            try:
                torch.cuda.set_device(local_rank)
                for _ in range(counts[i]):
                    outs[i] = fn(i)
            except Exception as e:  # surfaced after the join
                errs.append(e)

        barrier()
        ev0.record(main_stream)
        if workers == 1:
            work(0)
        else:
            ths = [threading.Thread(target=work, args=(i,)) for i in range(workers)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        if errs:
            raise errs[0]
        for st_ in streams:
            main_stream.wait_stream(st_)
        ev1.record(main_stream)
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1)), [o for o in outs if o is not None]

    for i in range(P):
        for _ in range(args.warmup):
            proof = step_resident(i)
    ref_vec = proof.to_vec().copy()

    # latency of one proof alone (one pipeline), then the timed throughput run with P in flight
    ms_lat, _ = timed(step_resident, 2, 1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for c in ctxs:
        c.stats_reset()
    ms_total, proofs = timed(step_resident, args.steps, P)
    sts = [c.stats() for c in ctxs]
    clocks = sampler.stop() if rank == 0 else None
    launches = sum_over_ranks(float(sum(x.launches_total for x in sts)))
    for pr in proofs:
        assert np.array_equal(pr.to_vec(), ref_vec), "non-deterministic proof"
    # per-proof host breakdown, averaged over the proofs this rank LED (sharded: followers neither hash nor wait on rounds)
    led = [x for x in sts if x.rounds > 0]
    rounds_per_proof = 92 * bn  # 91 cipher layers + the identity layer, bn rounds each
    n_led = max(1, int(round(sum(x.rounds for x in led) / max(1, rounds_per_proof)))) if led else 1
    brk = {"transcript_host": sum(x.transcript_ms for x in led) / n_led, "wait_device": sum(x.wait_ms for x in led) / n_led,
           "comm_host": sum(x.comm_ms for x in led) / n_led, "rounds": rounds_per_proof, "proofs_led_by_rank0": n_led if led else 0,
           "follower_asleep_rank0": sum(x.follow_wait_ms for x in sts) / max(1, args.steps)}

    # e2e through the host-buffer API
    for i in range(P):
        step_e2e(i)
    for c in ctxs:
        c.stats_reset()
    ms_e2e, proofs_e = timed(step_e2e, args.steps, P)
    sts_e = [c.stats() for c in ctxs]
    for pr in proofs_e:
        assert np.array_equal(pr.to_vec(), ref_vec), "host-buffer path and device-resident path disagree"
    h2d = sum_over_ranks(float(sum(x.h2d_bytes for x in sts_e))) / args.steps
    d2h = sum_over_ranks(float(sum(x.d2h_bytes for x in sts_e))) / args.steps
    out_sha_local = sha(out_hn[0])

    # roofline pass: one extra step with CUDA events around every kernel launch (on the launching stream)
    # (sharded: in lockstep mode for this one step -- in leader mode a round kernel's last block stays resident until the challenge
    # derived from its sums arrives, so launch-to-completion times would include the host transcript)
    if sharded and ctx.exchange_mode == "window":
        ctx.set_option(gkrb200.Context.OPT_TRANSCRIPT, 1)
    ctx.set_profiling(True)
    ctx.stats_reset()
    step_resident()
    sp = ctx.stats()
    ctx.set_profiling(False)
    if sharded and ctx.exchange_mode == "window" and "7=1" not in args.opt:
        ctx.set_option(gkrb200.Context.OPT_TRANSCRIPT, 0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    k_ms = sp.kernel_ms[2]
    k_launches = max(int(sp.launches[2]), 1)
    achieved_gbs = sp.bytes_round / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    imad_rate = max(ctx.microbench(0, 2000)[0], ctx.microbench(2, 2000)[0])  # G IMAD.WIDE.U32 / s (independent and carry-chained forms), measured live
    frmul_rate, _ = ctx.microbench(1, 1000)      # G Fr-mul/s of the library's own multiplier at full occupancy
    int_peak = max(imad_rate / 136.0, frmul_rate)  # 136 wide MACs per Montgomery product; the multiplier itself sustains the pipe best
    achieved_mul = sp.fr_mul_round / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    roofline = {"bound": "integer", "kernel": "k_round_cf (fold + factored round evaluation; K3+K4)",
                "achieved": achieved_mul, "peak": int_peak, "unit": "G Fr-mul/s", "frac": achieved_mul / int_peak if int_peak else None,
                "traffic": NCU_TRAFFIC.get(args.bn), "traffic_note": NCU_TRAFFIC_NOTE,
                "peak_source": "measured live: max(IMAD.WIDE.U32 rate / 136 wide multiply-adds per Montgomery product, the library multiplier's own microbenchmark); "
                               "no integer peak in MEASURED_PEAKS.json",
                "algorithmic_fr_mul_per_launch": sp.fr_mul_round / k_launches, "launches": k_launches, "avg_launch_us": k_ms * 1e3 / k_launches,
                "imad_wide_gmacs_measured": imad_rate, "macs_per_fr_mul": 136, "fr_mul_microbench_gmuls": frmul_rate,
                "share_of_step_kernel_time": k_ms / max(sum(sp.kernel_ms), 1e-9),
                "note": "integer-multiply pipe (IMAD.WIDE.U32 on fmaheavy) is the bound of this kernel: 13-17 product-equivalents per 128-384 B; the HBM view is in roofline_hbm"}
    roofline_hbm = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": sp.bytes_round / k_launches}
    kernels = {name: {"launches": int(sp.launches[i]), "ms": sp.kernel_ms[i]} for i, name in
               enumerate(["assign", "eq", "round", "fold", "multi_eq", "staging", "misc"])}

    # parity: the batch the cpu_baseline leg proves, proven here with the same seeds (sharded over all ranks when N > 1)
    parity = cpu_baseline = None
    if not args.no_cpu_baseline:
        pbn = args.cpu_bn
        pn = 1 << pbn
        pin = (synth_inputs(pn, args.seed), synth_inputs(pn, args.seed + 1), synth_inputs(pbn, args.seed + 2))
        pa = circuits[0].Assign(pin[0], pin[1], want_outputs=True)
        pvec = gkrb200.gkr.Prove(circuits[0], pa, pin[2]).to_vec().copy()
        gpu_out = pa.outputs
        if dist is not None and pa.n_local != pn:  # gather the strided output shards on rank 0
            parts = [torch.empty((pa.n_local, 4), dtype=torch.int64, device="cuda") for _ in range(world)] if rank == 0 else None
            dist.gather(torch.from_numpy(gpu_out.view(np.int64)).cuda(), parts, dst=0)
            if rank == 0:
                full = np.empty((pn, 4), dtype=np.uint64)
                for g in range(world):
                    full[g::world] = parts[g].cpu().numpy().view(np.uint64)
                gpu_out = full
        if rank == 0:
            v, secs, o93, ovec = cpu_port_prove(pbn, cores, args.seed, pin)
            parity = {"bn": pbn, "n_gpus": world, "sha256_gpu": sha(pvec), "sha256_oracle": sha(ovec), "equal": bool(np.array_equal(pvec, ovec)),
                      "hash_outputs_sha256_gpu": sha(gpu_out), "hash_outputs_sha256_oracle": sha(o93), "hash_outputs_equal": bool(np.array_equal(gpu_out, o93)),
                      "what": "GkrProofToVec image (Montgomery words) and a[93] of a 2^%d batch, seeds %s, against the CPU oracle (oracle/gkr_oracle.c)" % (pbn, [args.seed, args.seed + 1, args.seed + 2])}
            cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": "full Assign+Prove of a 2^%d-hash batch (%.1f s), C port of the reference Go prover (Go toolchain absent), %d threads" % (pbn, secs, cores),
                            **fr_mul_ns()}

    if rank == 0:
        ms_step = ms_total / args.steps
        n_step = n * (world if replicas else 1)  # hashes proven per step over all ranks
        line = {
            "metric": METRIC, "value": n_step / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak" if replicas else "strong", "vs_baseline": None, "dtype": "u32x8 (BN254 Fr, Montgomery)",
            "data": "synthetic", "config": workload_config(args, world, P, replicas, ctx.exchange_mode), "clocks": clocks,
            "e2e": {"value": n_step / (ms_e2e / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "returns": "a[93] (hash outputs, %d MiB%s) and the proof vector, into pinned host memory" % (
                        (n_local * 32) >> 20, " per rank" if world > 1 else "")},
            "gpu_launches": int(launches),
            "roofline": roofline, "roofline_hbm": roofline_hbm, "kernels_profile_step": kernels,
            "parity": parity, "proof_sha256": sha(ref_vec), "hash_outputs_sha256_rank0": out_sha_local,
            "pipeline": {"proofs_in_flight": P, "latency_ms_one_proof_alone": ms_lat / 2,
                         "note": "value/e2e time K proofs with P in flight (own context, stream and host thread each): the serial host transcript of one overlaps the device rounds of another"},
            "breakdown_ms_per_proof": brk,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    for c in ctxs:
        c.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_sumcheck_config(args, env):
    """BASELINE config 2: standalone sumcheck.Prove (sumcheck/prover.go:46-90) of eq * CipherGate over 2^bn-entry synthetic L/R
    tables on ONE B200 (sumcheck/prover_test.go:96-109 BenchmarkWithCipherGate style).  N > 1: independent replicas (the
    standalone API does not shard)."""
    import numpy as np
    import torch
    import gkrb200
    rank, local_rank, world = env["rank"], env["local_rank"], env["world"]
    bn, n = args.bn, 1 << args.bn
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = gkrb200.Context(device=local_rank, max_bn=bn, stream=stream.cuda_stream)
    L, R, q = synth_inputs(n, args.seed), synth_inputs(n, args.seed + 1), synth_inputs(bn, args.seed + 2).reshape(1, bn, 4)
    ark = synth_inputs(1, args.seed + 3)[0]
    gate = gkrb200.gates.CipherGate(ark)
    claim = np.zeros((1, 4), dtype=np.uint64)
    L_h, R_h = torch.from_numpy(L.view(np.int64)).pin_memory(), torch.from_numpy(R.view(np.int64)).pin_memory()
    L_d, R_d = L_h.cuda(), R_h.cuda()
    L_hn, R_hn = L_h.numpy().view(np.uint64), R_h.numpy().view(np.uint64)
    torch.cuda.synchronize()

    def run(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        env["barrier"]()
        ev0.record(stream)
        for _ in range(steps):
            out = fn()
        ev1.record(stream)
        env["barrier"]()
        return env["max_over_ranks"](ev0.elapsed_time(ev1)), out

    resident = lambda: gkrb200.sumcheck.ProveDevice(ctx, [L_d.data_ptr(), R_d.data_ptr()], q, claim, gate)
    hostbuf = lambda: gkrb200.sumcheck.Prove(ctx, [L_hn, R_hn], q, claim, gate)
    for _ in range(max(args.warmup, 3)):
        ref = resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.stats_reset()
    ms_total, out = run(resident, args.steps)
    st = ctx.stats()
    clocks = sampler.stop() if rank == 0 else None
    assert all(np.array_equal(a, b) for a, b in zip(out, ref)), "non-deterministic sumcheck proof"
    hostbuf()
    ctx.stats_reset()
    ms_e2e, out_e = run(hostbuf, args.steps)
    st_e = ctx.stats()
    assert all(np.array_equal(a, b) for a, b in zip(out_e, ref)), "host-buffer and device-resident paths disagree"
    ctx.set_profiling(True)
    ctx.stats_reset()
    resident()
    sp = ctx.stats()
    ctx.set_profiling(False)
    k_ms, k_launches = sp.kernel_ms[2], max(int(sp.launches[2]), 1)
    imad_rate = max(ctx.microbench(0, 2000)[0], ctx.microbench(2, 2000)[0])
    frmul_rate, _ = ctx.microbench(1, 1000)
    int_peak = max(imad_rate / 136.0, frmul_rate)
    achieved_mul = sp.fr_mul_round / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    parity = cpu_baseline = None
    cores = os.cpu_count() or 1
    if rank == 0 and not args.no_cpu_baseline:
        coracle = _coracle(cores)
        t0 = time.perf_counter()
        o_proof, o_chal, o_fin = coracle.sumcheck_prove([L, R], q, claim, coracle.GATE_CIPHER, ark)
        secs = time.perf_counter() - t0
        eq_all = bool(np.array_equal(ref[0], o_proof) and np.array_equal(ref[1], o_chal) and np.array_equal(ref[2], o_fin))
        parity = {"bn": bn, "sha256_gpu": sha(np.concatenate([ref[0].reshape(-1, 4), ref[1], ref[2]])),
                  "sha256_oracle": sha(np.concatenate([o_proof.reshape(-1, 4), o_chal, o_fin])), "equal": eq_all,
                  "what": "round polynomials, challenges and final claims of sumcheck.Prove against the CPU oracle"}
        cpu_baseline = {"value": n / secs, "unit": "entries/s", "cores": cores, "kind": "port",
                        "sample": "one standalone cipher-gate sumcheck.Prove over 2^%d-entry tables (%.1f s), C port of the Go prover, %d threads" % (bn, secs, cores), **fr_mul_ns()}
    if rank == 0:
        ms_step = ms_total / args.steps
        line = {
            "metric": "standalone sumcheck (eq * cipher gate) table entries/sec", "value": n * world / (ms_step * 1e-3), "unit": "entries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 (BN254 Fr, Montgomery)", "data": "synthetic",
            "config": {"workload": "standalone sumcheck.Prove of eq * CipherGate(L, R) over 2^%d-entry synthetic tables (%d rounds, 9 coefficients each)" % (bn, bn),
                       "baseline_config": CONFIG_NAME[2], "bn": bn, "entries_per_step": n * world, "parallelism": "single GPU, one sumcheck at a time" if world == 1 else "%d independent replicas" % world,
                       "l2": "inputs larger than L2? no: two %d MiB tables; every step re-reads them from HBM-resident copies after the previous step's folds overwrote L2" % ((32 << bn) >> 20),
                       "seeds": [args.seed, args.seed + 1, args.seed + 2, args.seed + 3]},
            "clocks": clocks,
            "e2e": {"value": n * world / (ms_e2e / args.steps * 1e-3), "unit": "entries/s", "h2d_bytes_per_step": st_e.h2d_bytes / args.steps * world,
                    "d2h_bytes_per_step": st_e.d2h_bytes / args.steps * world, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(env["sum_over_ranks"](float(st.launches_total))),
            "roofline": {"bound": "integer", "kernel": "k_round_cf", "achieved": achieved_mul, "peak": int_peak, "unit": "G Fr-mul/s",
                         "frac": achieved_mul / int_peak if int_peak else None, "traffic": None, "launches": k_launches, "avg_launch_us": k_ms * 1e3 / k_launches},
            "parity": parity, "cpu_baseline": cpu_baseline,
            "breakdown_ms_per_step": {"transcript_host": st.transcript_ms / args.steps, "wait_device": st.wait_ms / args.steps},
        }
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    ctx.close()
    if env["dist"] is not None:
        env["dist"].destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
