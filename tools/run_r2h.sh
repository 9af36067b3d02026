#!/bin/bash
# round 2, run H (1 GPU): parity suite with the constant-multiplier fold; default bench; A/B const-fold off; launch list
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/h_pytest.log 2>&1; echo "pytest done"; cat gpurun_out/h_pytest.log
timeout 600 python bench.py --steps 16 --warmup 3 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/h_bench.err
timeout 300 python bench.py --steps 8 --warmup 2 --no-cpu-baseline --opt 8=0 > gpurun_out/h_bench_nokfold.json 2> gpurun_out/h_bench_nokfold.err; echo "bench nokfold rc=$?"
timeout 300 python bench.py --steps 12 --warmup 2 --no-cpu-baseline --inflight 12 > gpurun_out/h_bench_p12.json 2> gpurun_out/h_bench_p12.err; echo "bench p12 rc=$?"
for f in h_bench h_bench_nokfold h_bench_p12; do python - $f <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/%s.json"%sys.argv[1]))
    k=d["kernels_profile_step"]
    print(sys.argv[1], "value %.2f M/s"%(d["value"]/1e6), "ms/step %.1f"%d["ms_per_step"], "e2e %.2f"%(d["e2e"]["value"]/1e6), "lat %.0f"%d["pipeline"]["latency_ms_one_proof_alone"], "assign %.1f round %.1f multi_eq %.1f eq %.1f"%(k["assign"]["ms"],k["round"]["ms"],k["multi_eq"]["ms"],k["eq"]["ms"]), "frac %.3f"%d["roofline"]["frac"], "parity", (d.get("parity") or {}).get("equal"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e: print(sys.argv[1], "failed", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/h_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --inflight 1 > gpurun_out/h_launches_bench.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/h_launches.csv > gpurun_out/h_launches.txt 2>&1; head -16 gpurun_out/h_launches.txt
