#!/bin/bash
# round 2, run A: parity suite (incl. the new word-for-word tests at 2^16/2^20/2^22) + A/B of the out-of-line vs inlined multiplier builds
set -u
mkdir -p gpurun_out
nproc > gpurun_out/a_nproc.txt; nvidia-smi -L >> gpurun_out/a_nproc.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/a_pytest.log
for V in "" inl; do
  export GKRB200_LIB_VARIANT=$V
  T=${V:-base}
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --inflight 1 > gpurun_out/a_bench1_$T.json 2> gpurun_out/a_bench1_$T.err; echo "bench1[$T] rc=$?"
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --inflight 8 > gpurun_out/a_bench8_$T.json 2> gpurun_out/a_bench8_$T.err; echo "bench8[$T] rc=$?"
  python - "$T" <<'PY'
import json,sys
f=sys.argv[1]
for w in ("1","8"):
    try:
        d=json.load(open("gpurun_out/a_bench%s_%s.json"%(w,f)))
        k=d["kernels_profile_step"]
        print(f, w, "value %.2f M/s"%(d["value"]/1e6), "ms/step %.1f"%d["ms_per_step"], "lat %.0f"%d["pipeline"]["latency_ms_one_proof_alone"], "assign %.1f round %.1f multi_eq %.1f eq %.1f"%(k["assign"]["ms"],k["round"]["ms"],k["multi_eq"]["ms"],k["eq"]["ms"]), "int frac %.3f"%d["roofline_int"]["frac"], "peak %.1f"%d["roofline_int"]["peak"])
    except Exception as e:
        print(f, w, "failed", e)
PY
done
