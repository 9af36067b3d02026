#!/bin/bash
# round 2, multi-GPU run (N = 4 or 8): sharded parity in every transcript mode, then the scaling benches.  Usage: gpurun --gpus N -- env N=N bash tools/run_r2n.sh
set -u
mkdir -p gpurun_out
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L > gpurun_out/n${N}_gpus.txt 2>&1; nproc >> gpurun_out/n${N}_gpus.txt
GKRB200_VERBOSE=1 timeout 600 $TR --master-port 29711 tools/multigpu_check.py 16 > gpurun_out/n${N}_check.log 2>&1; echo "check rc=$?"
grep -E "bn=|rror" gpurun_out/n${N}_check.log | tail -24
GKRB200_VERBOSE=1 timeout 300 $TR --master-port 29715 tools/multigpu_check.py 12 last > gpurun_out/n${N}_check_last.log 2>&1; echo "check(last leader) rc=$?"
grep -c "OK (bit-exact" gpurun_out/n${N}_check_last.log
run() { # name, args...
  local name=$1; shift
  timeout 600 $TR --master-port 29720 bench.py --gpus $N "$@" > gpurun_out/n${N}_bench_$name.json 2> gpurun_out/n${N}_bench_$name.err; echo "bench $name rc=$?"
  python - gpurun_out/n${N}_bench_$name.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("  value %.2f M/s"%(d["value"]/1e6), "ms/step %.1f"%d["ms_per_step"], "e2e %.2f"%(d["e2e"]["value"]/1e6), "h2d %.0f MB"%(d["e2e"]["h2d_bytes_per_step"]/1e6), "d2h %.0f MB"%(d["e2e"]["d2h_bytes_per_step"]/1e6),
          "P", d["pipeline"]["proofs_in_flight"], "lat %.0f"%d["pipeline"]["latency_ms_one_proof_alone"], d["breakdown_ms_per_proof"], "parity", (d.get("parity") or {}).get("equal"), d["proof_sha256"][:16])
except Exception as e: print("  failed", e)
PY
}
run leader --steps 16 --warmup 2
run leader_p16 --steps 32 --warmup 2 --inflight 16 --no-cpu-baseline
run lockstep --steps 16 --warmup 2 --opt 7=1 --inflight $(( $(nproc) / N )) --no-cpu-baseline
run replicas --steps 16 --warmup 2 --mode replicas --no-cpu-baseline
