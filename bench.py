#!/usr/bin/env python3
"""bench.py -- proven MiMC hashes/sec (bit-exact GKR proof) on N B200s.

One "step" = one full pass of the hot path over one batch: Circuit.Assign (circuit/assignment.go:12) +
gkr.Prove (gkr/prover.go:21) of a 2^bn-hash batch through the C ABI of libgkrb200.so.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
  python bench.py --impl reference ...      # the CPU port of the reference's Go prover on the host cores

`value`  : whole-job hashes/s with the inputs already resident in HBM when the timed region starts.
`e2e`    : same metric through the host-buffer API (gkrb200_mimc_assign + gkrb200_gkr_prove_mimc), H2D of the
           inputs and D2H of every result inside the timed region.
`roofline`: the dominant kernel (k_round: fold + round evaluation) timed live with CUDA events around every
           launch of one extra step; algorithmic bytes / that time vs the measured HBM peak, plus `roofline_int`
           (field multiplications / that time vs the integer-pipe peak measured by the library's microbenchmark).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "gkr-mimc_b200"))

_JSON_OUT = sys.stdout
METRIC = "proven MiMC hashes/sec (bit-exact GKR proof)"
UNIT = "hashes/s"
Q3 = 0x30644E72E131A029  # top limb of q: any element with top limb < Q3 is canonical


def synth_inputs(n, seed):
    """Deterministic pseudo-random canonical field elements (Go layout). Seeds are recorded in the JSON line."""
    import numpy as np
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    a[:, :3] |= rng.integers(0, 2, size=(n, 3), dtype=np.uint64) << np.uint64(63)
    a[:, 3] %= np.uint64(Q3)
    return a


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


def cpu_port_hashes_per_s(bn, threads, seed, reps=1):
    """Times the CPU port of the reference prover (oracle/, test infrastructure) on a 2^bn batch."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import coracle
    coracle.build()
    coracle.set_threads(threads)
    n = 1 << bn
    key, msg, qp = synth_inputs(n, seed), synth_inputs(n, seed + 1), synth_inputs(bn, seed + 2)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        coracle.assign_and_prove_mimc(key, msg, qp)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return n / best, best


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The Go toolchain
    does not exist in this image, so this is the C port of the Go prover (oracle/gkr_oracle.c, same task
    decomposition, one worker per core)."""
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    bn = args.ref_bn
    for _ in range(args.warmup):
        cpu_port_hashes_per_s(min(bn, 12), cores, args.seed)
    t0 = time.perf_counter()
    n_done = 0
    for _ in range(args.steps):
        cpu_port_hashes_per_s(bn, cores, args.seed)
        n_done += 1 << bn
    dt = time.perf_counter() - t0
    value = n_done / dt
    sample = "each step = full Assign+Prove of a 2^%d-hash batch (bounded sample of the 2^%d workload), C port of the Go prover, %d threads" % (bn, args.bn, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64x4 (BN254 Fr, Montgomery)",
        "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    return 0


# dram__bytes_read.sum + dram__bytes_write.sum of the largest k_round_cf launches (ncu --set full, profiles/r1_ncu_k_round_cf_final_bn22.txt):
# round 0 (2^21 pairs, no fold) 272.7 MB (7 sums) / 273.8 MB (8 sums) vs 268.4 MB algorithmic; round 1 (2^20 pairs, fold) 377.2 MB vs 402.7 MB algorithmic
NCU_TRAFFIC = {22: 272.7e6}
NCU_TRAFFIC_NOTE = "ncu capture of the round-0 launch of one layer (2^21 pairs: 268.4 MB algorithmic); `achieved` averages all 1564 launches of a proof"


def workload_config(args, world, P=1, replicas=False, exchange=None):
    xdesc = {"window": "round sums published by the round kernels into a host-shared mapped window read by every rank (no collective on the round path)",
             "nccl": "NCCL all-gather of the round sums"}.get(exchange, "round sums exchanged every round")
    return {
        "workload": "full MiMC GKR proof (Circuit.Assign + gkr.Prove, 94 layers, transcript bit-exact) of a 2^%d-hash batch over BN254 Fr" % args.bn,
        "bn": args.bn, "hashes_per_step": (1 << args.bn) * (world if replicas else 1), "proof_elements": 1006 * args.bn + 183,
        "parallelism": "single GPU, %d proofs in flight" % P if world == 1 else (
            "replicas: each of the %d GPUs proves its own 2^%d batches, no exchange, %d proofs in flight per GPU" % (world, args.bn, P) if replicas else
            "batch sharded on low address bits over %d GPUs, %s, %d proofs in flight" % (world, xdesc, P)),
        "l2": "inputs larger than L2: 93 layer tables of %d MiB each per proof" % ((32 << args.bn) >> 20),
        "seeds": [args.seed, args.seed + 1, args.seed + 2],
        **({"options": args.opt} if args.opt else {}),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bn", type=int, default=22, help="log2 of the batch (BASELINE.json quotes the metric on 2^22)")
    ap.add_argument("--ref-bn", type=int, default=16, help="batch of one --impl reference step (bounded sample)")
    ap.add_argument("--cpu-bn", type=int, default=20, help="batch of the cpu_baseline sample")
    ap.add_argument("--seed", type=int, default=0x6B6B72)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=0, help="proofs in flight per GPU (0 = auto: up to 8, bounded by --steps and by the host cores per rank)")
    ap.add_argument("--mode", default="sharded", choices=["sharded", "replicas"],
                    help="N > 1 GPUs: 'sharded' = every 2^bn batch is split over all N GPUs (one proof, round sums exchanged; the north-star "
                         "configuration, strong scaling); 'replicas' = every GPU proves its own 2^bn batches (no exchange, weak scaling)")
    ap.add_argument("--opt", action="append", default=[], metavar="ID=VALUE", help="gkrb200_set_option passthrough (tuning experiments), repeatable")
    args = ap.parse_args()

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

    n, bn = 1 << args.bn, args.bn
    # P proofs in flight: each has its own context (arena + stream) and its own host thread, so the serial host
    # transcript of one proof (MiMC challenges, ~130 ms per 2^22 proof) overlaps the device rounds of another.
    replicas = world > 1 and args.mode == "replicas"
    sharded = world > 1 and not replicas
    cores = os.cpu_count() or 1
    # every pipeline keeps one host thread busy (transcript or spinning on the result slot): stay within the cores a rank can have
    # (measured on one B200, 2^22: P = 3 -> 174 ms/proof, 4 -> 165, 6 -> 155, 8 -> 151, 12 -> 147.5: more proofs in flight let the
    # latency-bound small rounds of one proof run beside the big rounds of another; 8 x 12.4 GB of arenas fit in 180 GB)
    P = args.inflight if args.inflight > 0 else max(1, min(8, args.steps, cores // world))
    main_stream = torch.cuda.Stream()
    streams = [torch.cuda.Stream() for _ in range(P)]  # the library launches on these; events are recorded on main_stream after joining them
    torch.cuda.set_stream(main_stream)
    ctxs = [gkrb200.Context(device=local_rank, max_bn=bn, stream=st_.cuda_stream) for st_ in streams]
    ctx = ctxs[0]
    if sharded:
        for c in ctxs:  # one communicator per pipeline, created in the same order on every rank
            uid = [gkrb200.Context.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            c.comm_init(rank, world, uid[0])
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
    torch.cuda.synchronize()

    def step_resident(i=0):
        a = circuits[i].AssignDevice(key_d.data_ptr(), msg_d.data_ptr(), n)
        return gkrb200.gkr.Prove(circuits[i], a, q_np)

    def step_e2e(i=0):
        a = circuits[i].Assign(key_hn, msg_hn)
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
    st = sts[0]
    clocks = sampler.stop() if rank == 0 else None
    launches = sum_over_ranks(float(sum(x.launches_total for x in sts)))
    for pr in proofs:
        assert np.array_equal(pr.to_vec(), ref_vec), "non-deterministic proof"

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
    circuit = circuits[0]

    # roofline pass: one extra step with CUDA events around every kernel launch (on the launching stream)
    ctx.set_profiling(True)
    ctx.stats_reset()
    step_resident()
    sp = ctx.stats()
    ctx.set_profiling(False)
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
    roofline = {"bound": "hbm", "kernel": "k_round_cf (fold + factored round evaluation; K3+K4)", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved_gbs / hbm_peak, "traffic": NCU_TRAFFIC.get(args.bn), "traffic_note": NCU_TRAFFIC_NOTE, "peak_source": peak_src, "launches": k_launches,
                "avg_launch_us": k_ms * 1e3 / k_launches, "algorithmic_bytes_per_launch": sp.bytes_round / k_launches,
                "note": "k_round_cf is integer-pipe bound (19-23 Fr-mul per 128-384 B), not HBM bound: see roofline_int"}
    roofline_int = {"bound": "integer multiply pipe (IMAD.WIDE.U32 on fmaheavy: 32x32->64 multiply-add, measured at ~0.45x the 32-bit IMAD rate)", "achieved": achieved_mul, "peak": int_peak, "unit": "G Fr-mul/s",
                    "frac": achieved_mul / int_peak if int_peak else None, "imad_wide_gmacs_measured": imad_rate, "macs_per_fr_mul": 136,
                    "fr_mul_microbench_gmuls": frmul_rate, "frac_of_fr_mul_microbench": achieved_mul / frmul_rate if frmul_rate else None,
                    "share_of_step_kernel_time": k_ms / max(sum(sp.kernel_ms), 1e-9)}
    kernels = {name: {"launches": int(sp.launches[i]), "ms": sp.kernel_ms[i]} for i, name in
               enumerate(["assign", "eq", "round", "fold", "multi_eq", "staging", "misc"])}

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, secs = cpu_port_hashes_per_s(args.cpu_bn, cores, args.seed)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "full Assign+Prove of a 2^%d-hash batch (%.1f s), C port of the reference Go prover (Go toolchain absent), %d threads" % (args.cpu_bn, secs, cores)}

    if rank == 0:
        ms_step = ms_total / args.steps
        n_step = n * (world if replicas else 1)  # hashes proven per step over all ranks
        line = {
            "metric": METRIC, "value": n_step / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak" if replicas else "strong", "vs_baseline": None, "dtype": "u32x8 (BN254 Fr, Montgomery)",
            "data": "synthetic", "config": workload_config(args, world, P, replicas, ctx.exchange_mode), "clocks": clocks,
            "e2e": {"value": n_step / (ms_e2e / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": roofline, "roofline_int": roofline_int, "kernels_profile_step": kernels,
            "pipeline": {"proofs_in_flight": P, "latency_ms_one_proof_alone": ms_lat / 2,
                         "note": "value/e2e time K proofs with P in flight (own context, stream and host thread each): the serial host transcript of one overlaps the device rounds of another"},
            "breakdown_ms_per_proof_pipeline0": {"transcript_host": st.transcript_ms / max(1, (args.steps + P - 1) // P), "wait_device": st.wait_ms / max(1, (args.steps + P - 1) // P),
                                      "comm_host": st.comm_ms / max(1, (args.steps + P - 1) // P), "rounds": int(st.rounds // max(1, (args.steps + P - 1) // P))},
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    for c in ctxs:
        c.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
