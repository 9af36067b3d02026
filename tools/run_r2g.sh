#!/bin/bash
# round 2, run G (1 GPU): parity suite after the fold / identity-round / eq changes; BASELINE configs 1-4; HBM-side kernels + ncu DRAM counters
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/g_pytest.log
timeout 300 python tools/gpu_hbm_kernels.py 22 > gpurun_out/g_hbm.txt 2>&1; cat gpurun_out/g_hbm.txt
for K in k_fold k_eq_expand "k_round<"; do
  N=$(echo $K | tr -cd 'a-z_')
  timeout 300 ncu --set full --clock-control none -k "regex:$K" -c 1 -o gpurun_out/g_ncu_$N -f python tools/gpu_hbm_kernels.py 22 1 > gpurun_out/g_ncu_$N.log 2>&1; echo "ncu $N rc=$?"
  python tools/ncu_summary.py gpurun_out/g_ncu_$N.ncu-rep > gpurun_out/g_ncu_$N.txt 2>&1
  grep -E "==|time_duration|dram__bytes|dram_throughput|fmaheavy" gpurun_out/g_ncu_$N.txt
done
for C in 1 2 3 4; do
  timeout 600 python bench.py --config $C --steps 8 --warmup 3 > gpurun_out/g_bench_config$C.json 2> gpurun_out/g_bench_config$C.err; echo "bench config $C rc=$?"; tail -2 gpurun_out/g_bench_config$C.err
  python - $C <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/g_bench_config%s.json"%sys.argv[1]))
    print(d["metric"][:40], "value %.3g %s"%(d["value"], d["unit"]), "ms/step %.2f"%d["ms_per_step"], "e2e %.3g"%d["e2e"]["value"], "frac", round(d["roofline"]["frac"],3), "parity", (d.get("parity") or {}).get("equal"), "cpu %.3g"%(d["cpu_baseline"] or {}).get("value",0))
except Exception as e: print("failed", e)
PY
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/g_bench_reference.json 2> gpurun_out/g_bench_reference.err; echo "reference rc=$?"; cut -c1-600 gpurun_out/g_bench_reference.json
