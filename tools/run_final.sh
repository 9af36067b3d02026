#!/bin/bash
# Round-end evidence on one B200: parity suite, default bench (both arms), launch list, ncu captures of the top kernels.
# Usage: gpurun --timeout 900 -- bash tools/run_final.sh
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/f_pytest.log
timeout 600 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/f_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err; echo "reference rc=$?"
# launch list of one un-pipelined step (ncu serialises and runs cold: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/f_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --inflight 1 > gpurun_out/f_launches_bench.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/f_launches.csv > gpurun_out/f_launches.txt 2>&1; head -14 gpurun_out/f_launches.txt
# full captures: first round-0 and first fold launch of k_round_cf (2^21 / 2^20 pairs), the multi-claim eq kernel, the assign kernel
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_round_cf -c 2 -o gpurun_out/f_ncu_round -f python tools/gpu_prove_once.py 22 > gpurun_out/f_ncu_round.log 2>&1; echo "ncu round rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_eq_expand -c 1 -o gpurun_out/f_ncu_eq -f python tools/gpu_prove_once.py 22 > gpurun_out/f_ncu_eq.log 2>&1; echo "ncu eq rc=$?"
ls -la gpurun_out/*.ncu-rep
python - <<'PY'
import json
d=json.load(open("gpurun_out/f_bench.json"))
print("value %.2f M/s"%(d["value"]/1e6), "ms/step %.1f"%d["ms_per_step"], "e2e %.2f"%(d["e2e"]["value"]/1e6), "P", d["pipeline"]["proofs_in_flight"], "cpu", d["cpu_baseline"]["value"], "roofline", d["roofline"]["frac"], d["roofline_int"]["frac"])
PY
