#!/bin/bash
# proofs-in-flight sweep on one GPU.  Usage: gpurun -- bash tools/run_sweep.sh "8 10 12"
set -u
mkdir -p gpurun_out
for P in ${1:-8 10 12}; do
  timeout 300 python bench.py --inflight $P --steps $((2*P)) --warmup 3 --no-cpu-baseline > gpurun_out/n1_bench_p$P.json 2> gpurun_out/n1_bench_p$P.err; echo "bench P=$P rc=$?"; tail -2 gpurun_out/n1_bench_p$P.err
done
timeout 300 python bench.py --inflight 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/n1_bench_p5k5.json 2> gpurun_out/n1_bench_p5k5.err; echo "bench P=5 K=5 rc=$?"
for f in ${1:-8 10 12} 5k5; do python - "p$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open("gpurun_out/n1_bench_%s.json"%f))
    print(f, "value %.2f M/s"%(d["value"]/1e6), "ms/step %.1f"%d["ms_per_step"], "e2e %.2f"%(d["e2e"]["value"]/1e6), "steps",d["steps"], d["pipeline"]["proofs_in_flight"], "lat %.0f"%d["pipeline"]["latency_ms_one_proof_alone"], d["breakdown_ms_per_proof_pipeline0"])
except Exception as e:
    print(f, "failed", e)
PY
done
