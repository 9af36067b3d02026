#!/bin/bash
# round 2, run C: inline (opt 6=0) vs call (6=1e9) vs only >= 2^20 pairs inline (6=1048576) at 2, 4, 8 proofs in flight
set -u
mkdir -p gpurun_out
for P in 2 4 8; do
for T in 0 1048576 1000000000; do
  timeout 300 python bench.py --steps 8 --warmup 2 --no-cpu-baseline --inflight $P --opt 6=$T > gpurun_out/c_bench${P}_$T.json 2> gpurun_out/c_bench${P}_$T.err; echo "bench P=$P [$T] rc=$?"
  python - "$P" "$T" <<'PY'
import json,sys
p,f=sys.argv[1],sys.argv[2]
try:
    d=json.load(open("gpurun_out/c_bench%s_%s.json"%(p,f)))
    k=d["kernels_profile_step"]
    print(p, f, "value %.2f M/s"%(d["value"]/1e6), "ms/step %.1f"%d["ms_per_step"], "e2e %.2f"%(d["e2e"]["value"]/1e6), "lat %.0f"%d["pipeline"]["latency_ms_one_proof_alone"], "round %.1f"%(k["round"]["ms"]), "frac %.3f"%d["roofline"]["frac"])
except Exception as e:
    print(p, f, "failed", e)
PY
done
done
