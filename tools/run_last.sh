#!/bin/bash
# short 1-GPU confirmation: whole GPU suite (no -x: report every failure), smoke(), then a short bench if time is left
set -u
mkdir -p gpurun_out
( time timeout 120 python -m pytest tests -m gpu -q ) > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/g_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/g_smoke.log
timeout 80 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/g_bench.err; head -c 400 gpurun_out/g_bench.json
