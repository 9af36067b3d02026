#!/bin/bash
# round 2, run F (1 GPU): block-size / occupancy variants of the inlined round kernel; ncu of the default build; HBM-bound kernels
set -u
mkdir -p gpurun_out
for V in "" b128x3 b160x3 b192x2; do
  export GKRB200_LIB_VARIANT=$V
  T=${V:-b256x2}
  for P in 1 8; do
  timeout 300 python bench.py --steps 8 --warmup 2 --no-cpu-baseline --inflight $P > gpurun_out/f_bench${P}_$T.json 2> gpurun_out/f_bench${P}_$T.err; echo "bench[$T] P=$P rc=$?"
  python - "$P" "$T" <<'PY'
import json,sys
p,f=sys.argv[1],sys.argv[2]
try:
    d=json.load(open("gpurun_out/f_bench%s_%s.json"%(p,f)))
    k=d["kernels_profile_step"]
    print(p, f, "value %.2f M/s"%(d["value"]/1e6), "ms/step %.1f"%d["ms_per_step"], "lat %.0f"%d["pipeline"]["latency_ms_one_proof_alone"], "assign %.1f round %.1f"%(k["assign"]["ms"], k["round"]["ms"]), "frac %.3f"%d["roofline"]["frac"])
except Exception as e:
    print(p, f, "failed", e)
PY
  done
done
unset GKRB200_LIB_VARIANT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_round_cf -c 2 -o gpurun_out/f_ncu_round -f python tools/gpu_prove_once.py 22 > gpurun_out/f_ncu_round.log 2>&1; echo "ncu round rc=$?"
python tools/ncu_summary.py gpurun_out/f_ncu_round.ncu-rep > gpurun_out/f_ncu_round.txt 2>&1; grep -E "==|time_duration|fmaheavy|warps_active|registers_per_thread |stalled_(wait|dispatch|math|no_inst|long)|inst_executed.sum|local" gpurun_out/f_ncu_round.txt
timeout 300 python tools/gpu_hbm_kernels.py 22 > gpurun_out/f_hbm.txt 2>&1; cat gpurun_out/f_hbm.txt
