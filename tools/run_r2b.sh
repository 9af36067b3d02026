#!/bin/bash
# round 2, run B: ncu capture of the inlined k_round_cf (layer 93 round 0 and round 1 at 2^22) + A/B of the inline threshold with 1 and 8 proofs in flight
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_round_cf -c 2 -o gpurun_out/b_ncu_round -f python tools/gpu_prove_once.py 22 > gpurun_out/b_ncu_round.log 2>&1; echo "ncu round rc=$?"
python tools/ncu_summary.py gpurun_out/b_ncu_round.ncu-rep > gpurun_out/b_ncu_round.txt 2>&1; head -70 gpurun_out/b_ncu_round.txt
ncu -i gpurun_out/b_ncu_round.ncu-rep --page source --csv > gpurun_out/b_ncu_round_source.csv 2>/dev/null
for T in 0 65536 1000000000; do
  timeout 300 python bench.py --steps 8 --warmup 2 --no-cpu-baseline --inflight 8 --opt 6=$T > gpurun_out/b_bench8_$T.json 2> gpurun_out/b_bench8_$T.err; echo "bench8[$T] rc=$?"
  python - "$T" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open("gpurun_out/b_bench8_%s.json"%f))
    k=d["kernels_profile_step"]
    print(f, "value %.2f M/s"%(d["value"]/1e6), "ms/step %.1f"%d["ms_per_step"], "lat %.0f"%d["pipeline"]["latency_ms_one_proof_alone"], "assign %.1f round %.1f multi_eq %.1f eq %.1f"%(k["assign"]["ms"],k["round"]["ms"],k["multi_eq"]["ms"],k["eq"]["ms"]), "int frac %.3f"%d["roofline_int"]["frac"])
except Exception as e:
    print(f, "failed", e)
PY
done
