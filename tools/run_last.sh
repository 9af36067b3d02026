#!/bin/bash
# short 1-GPU A/B: default bench first (the measurement), then the prover parity subset of the GPU suite
set -u
mkdir -p gpurun_out
timeout 45 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/k_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/k_bench.json"))
k=d["kernels_profile_step"]
print("value %.2f M/s"%(d["value"]/1e6), "ms/step %.1f"%d["ms_per_step"], "e2e %.2f"%(d["e2e"]["value"]/1e6), "lat %.0f"%d["pipeline"]["latency_ms_one_proof_alone"], {n: round(v["ms"],2) for n,v in k.items()}, "int frac %.3f"%d["roofline_int"]["frac"])
PY
( time timeout 40 python -m pytest tests -m gpu -q -x -k "gkr_prove_matches or sumcheck_cipher_gate or factored or generic_and_factored or standalone_2pow20 or multi_identity" ) > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/k_pytest.log
