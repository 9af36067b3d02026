#!/bin/bash
# quick 1-GPU check: parity suite + default bench.  Usage: gpurun -- bash tools/run_quick.sh [bench args]
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/q_pytest.log
timeout 400 python bench.py --no-cpu-baseline "$@" > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/q_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/q_bench.json"))
k=d["kernels_profile_step"]
print("value %.2f M/s"%(d["value"]/1e6), "ms/step %.1f"%d["ms_per_step"], "e2e %.2f"%(d["e2e"]["value"]/1e6), "P", d["pipeline"]["proofs_in_flight"], "lat %.0f"%d["pipeline"]["latency_ms_one_proof_alone"], {n: round(v["ms"],2) for n,v in k.items()}, "int frac %.3f"%d["roofline_int"]["frac"])
PY
