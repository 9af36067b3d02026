#!/bin/bash
# 1-GPU validation: smoke, full GPU parity suite, default bench, proofs-in-flight sweep.  Usage: gpurun -- bash tools/run_n1.sh
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n1_gpus.txt 2>&1; nproc >> gpurun_out/n1_gpus.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/n1_smoke.log 2>&1; echo "smoke rc=$?"
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/n1_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/n1_pytest.log
timeout 400 python bench.py > gpurun_out/n1_bench_default.json 2> gpurun_out/n1_bench_default.err; echo "bench rc=$?"
for P in 4 5 6; do
  timeout 300 python bench.py --inflight $P --steps $((2*P)) --warmup 3 --no-cpu-baseline > gpurun_out/n1_bench_p$P.json 2> gpurun_out/n1_bench_p$P.err; echo "bench P=$P rc=$?"
done
timeout 300 python bench.py --inflight 3 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/n1_bench_p3.json 2> gpurun_out/n1_bench_p3.err; echo "bench P=3 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/n1_bench_reference.json 2> gpurun_out/n1_bench_reference.err; echo "reference rc=$?"
for f in default p3 p4 p5 p6; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open("gpurun_out/n1_bench_%s.json"%f))
    print(f, "value %.2f M/s"%(d["value"]/1e6), "ms/step %.1f"%d["ms_per_step"], "e2e %.2f"%(d["e2e"]["value"]/1e6), "steps",d["steps"], d["pipeline"]["proofs_in_flight"], "lat %.0f"%d["pipeline"]["latency_ms_one_proof_alone"], "int frac %.3f"%d["roofline_int"]["frac"])
except Exception as e:
    print(f, "failed", e)
PY
done
