#!/bin/bash
# round 2, run D (2 GPUs): full parity suite on the refactored library, sharded parity in all transcript modes, short benches
set -u
mkdir -p gpurun_out
N=${N:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L > gpurun_out/d_gpus.txt 2>&1; nproc >> gpurun_out/d_gpus.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/d_pytest.log
GKRB200_VERBOSE=1 timeout 600 $TR --master-port 29711 tools/multigpu_check.py 16 > gpurun_out/d_check.log 2>&1; echo "check rc=$?"
grep -E "bn=|Error|error" gpurun_out/d_check.log | tail -30
GKRB200_VERBOSE=1 timeout 600 $TR --master-port 29715 tools/multigpu_check.py 13 last > gpurun_out/d_check_last.log 2>&1; echo "check(last leader) rc=$?"
grep -E "bn=|Error|error" gpurun_out/d_check_last.log | tail -8
timeout 600 $TR --master-port 29712 bench.py --gpus $N --steps 8 --warmup 2 > gpurun_out/d_bench_leader.json 2> gpurun_out/d_bench_leader.err; echo "bench leader rc=$?"
timeout 300 $TR --master-port 29713 bench.py --gpus $N --steps 8 --warmup 2 --no-cpu-baseline --opt 7=1 --inflight 8 > gpurun_out/d_bench_lockstep.json 2> gpurun_out/d_bench_lockstep.err; echo "bench lockstep rc=$?"
for f in leader lockstep; do echo "== $f"; tail -3 gpurun_out/d_bench_$f.err; python - $f <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/d_bench_%s.json"%sys.argv[1]))
    print("value %.2f M/s"%(d["value"]/1e6), "ms/step %.1f"%d["ms_per_step"], "e2e %.2f"%(d["e2e"]["value"]/1e6), "h2d %.0f MB"%(d["e2e"]["h2d_bytes_per_step"]/1e6), "d2h %.0f MB"%(d["e2e"]["d2h_bytes_per_step"]/1e6), "P", d["pipeline"], d["breakdown_ms_per_proof"], "parity", d["parity"], d["proof_sha256"][:16])
except Exception as e: print("failed", e)
PY
done
