#!/bin/bash
# A/B of the multiplier builds on one GPU: parity suite, multiplier microbenchmarks, bench.  Usage: gpurun -- bash tools/run_ab.sh
set -u
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/ab_micro.txt 2>&1
import sys; sys.path.insert(0, "gkr-mimc_b200")
import gkrb200
ctx = gkrb200.Context(0, 10)
for kind, name in ((1, "inlined fr_mul"), (4, "out-of-line fr_mulc")):
    print("%-20s full occupancy: %.1f G/s" % (name, ctx.microbench(kind, 1000)[0]))
    for w in (8, 12, 16):
        print("%-20s %2d warps/SM   : %.1f G/s" % (name, w, ctx.microbench(kind | (w << 8), 1000)[0]))
ctx.close()
PY
cat gpurun_out/ab_micro.txt
for V in ""; do  # round 1 compared the kara / mixed builds here (tools/exp/fr_kara.cuh); make variant NAME=.. DEFS=.. builds others
  export GKRB200_LIB_VARIANT=$V
  T=${V:-kara}
  ( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/ab_pytest_$T.log 2>&1; echo "pytest[$T] rc=$?"; tail -2 gpurun_out/ab_pytest_$T.log
  timeout 300 python bench.py --inflight 8 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_$T.json 2> gpurun_out/ab_bench_$T.err; echo "bench[$T] rc=$?"
  python - "$T" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open("gpurun_out/ab_bench_%s.json"%f))
    k=d["kernels_profile_step"]
    print(f, "value %.2f M/s"%(d["value"]/1e6), "ms/step %.1f"%d["ms_per_step"], "lat %.0f"%d["pipeline"]["latency_ms_one_proof_alone"], "assign %.1f round %.1f multi_eq %.1f"%(k["assign"]["ms"],k["round"]["ms"],k["multi_eq"]["ms"]), "int frac %.3f"%d["roofline_int"]["frac"], "peak %.1f"%d["roofline_int"]["peak"])
except Exception as e:
    print(f, "failed", e)
PY
done
