#!/bin/bash
# round 2, multi-GPU run 2: leader mode with more proofs in flight (N = 4 or 8).  Usage: gpurun --gpus N -- env N=N bash tools/run_r2p.sh
set -u
mkdir -p gpurun_out
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nproc > gpurun_out/p${N}_nproc.txt
if [ "$N" = "4" ]; then
  GKRB200_VERBOSE=1 timeout 600 $TR --master-port 29711 tools/multigpu_check.py 16 > gpurun_out/p${N}_check.log 2>&1; echo "check rc=$?"
  grep -c "OK (bit-exact" gpurun_out/p${N}_check.log; grep -E "MISMATCH|rror" gpurun_out/p${N}_check.log | head -5
fi
run() { # name, args...
  local name=$1; shift
  timeout 600 $TR --master-port 29720 bench.py --gpus $N "$@" > gpurun_out/p${N}_bench_$name.json 2> gpurun_out/p${N}_bench_$name.err; echo "bench $name rc=$?"
  python - gpurun_out/p${N}_bench_$name.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("  value %.2f M/s"%(d["value"]/1e6), "ms/step %.1f"%d["ms_per_step"], "e2e %.2f"%(d["e2e"]["value"]/1e6), "h2d %.0f MB"%(d["e2e"]["h2d_bytes_per_step"]/1e6), "d2h %.0f MB"%(d["e2e"]["d2h_bytes_per_step"]/1e6),
          "P", d["pipeline"]["proofs_in_flight"], "lat %.0f"%d["pipeline"]["latency_ms_one_proof_alone"], d["breakdown_ms_per_proof"], "parity", (d.get("parity") or {}).get("equal"), d["proof_sha256"][:16])
except Exception as e: print("  failed", e)
PY
  tail -2 gpurun_out/p${N}_bench_$name.err | cut -c1-300
}
if [ "$N" = "8" ]; then
  run leader_auto --steps 48 --warmup 2
  run leader_p16 --steps 32 --warmup 2 --inflight 16 --no-cpu-baseline
else
  run leader_auto --steps 24 --warmup 2
  run replicas --steps 16 --warmup 2 --mode replicas --no-cpu-baseline
fi
