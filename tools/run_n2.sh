#!/bin/bash
# 2-GPU validation: sharded parity (both exchange paths), then short sharded / replicas benches.  Usage: gpurun --gpus 2 -- bash tools/run_n2.sh
set -u
mkdir -p gpurun_out
N=${N:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L > gpurun_out/n2_gpus.txt 2>&1; nproc >> gpurun_out/n2_gpus.txt; df -h /dev/shm >> gpurun_out/n2_gpus.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/n2_smoke.log 2>&1; echo "smoke rc=$?"
GKRB200_VERBOSE=1 timeout 400 $TR --master-port 29711 tools/multigpu_check.py 14 > gpurun_out/n2_check.log 2>&1; echo "check rc=$?"
tail -20 gpurun_out/n2_check.log
timeout 300 $TR --master-port 29712 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/n2_bench_window.json 2> gpurun_out/n2_bench_window.err; echo "bench window rc=$?"
timeout 300 $TR --master-port 29713 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline --opt 5=1 > gpurun_out/n2_bench_nccl.json 2> gpurun_out/n2_bench_nccl.err; echo "bench nccl rc=$?"
timeout 300 $TR --master-port 29714 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu-baseline --mode replicas > gpurun_out/n2_bench_replicas.json 2> gpurun_out/n2_bench_replicas.err; echo "bench replicas rc=$?"
for f in window nccl replicas; do echo "== $f"; cut -c1-900 gpurun_out/n2_bench_$f.json; tail -3 gpurun_out/n2_bench_$f.err; done
