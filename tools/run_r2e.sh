#!/bin/bash
# round 2, run E (1 GPU): parity suite on the refactored library; inline(128 regs) vs call at 1 and 8 proofs in flight
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/e_pytest.log
for P in 1 8; do
for T in 0 1000000000; do
  timeout 300 python bench.py --steps 8 --warmup 2 --no-cpu-baseline --inflight $P --opt 6=$T > gpurun_out/e_bench${P}_$T.json 2> gpurun_out/e_bench${P}_$T.err; echo "bench P=$P [$T] rc=$?"; tail -2 gpurun_out/e_bench${P}_$T.err
  python - "$P" "$T" <<'PY'
import json,sys
p,f=sys.argv[1],sys.argv[2]
try:
    d=json.load(open("gpurun_out/e_bench%s_%s.json"%(p,f)))
    k=d["kernels_profile_step"]
    print(p, f, "value %.2f M/s"%(d["value"]/1e6), "ms/step %.1f"%d["ms_per_step"], "e2e %.2f"%(d["e2e"]["value"]/1e6), "lat %.0f"%d["pipeline"]["latency_ms_one_proof_alone"], "round %.1f"%(k["round"]["ms"]), "frac %.3f"%d["roofline"]["frac"], "d2h %.0f MB"%(d["e2e"]["d2h_bytes_per_step"]/1e6))
except Exception as e:
    print(p, f, "failed", e)
PY
done
done
