"""The HBM-bound kernels at BASELINE size (2^22 entries), timed with CUDA events on the launching stream (library profiling) and
reported against the measured HBM peak: k_fold (K4), single-claim k_eq_expand (K2), k_round<identity> (layer 2's round kernel).
Run plain for the GB/s figures; run under `ncu --set full -k regex:<kernel> -c 1` for the DRAM byte counters."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gkr-mimc_b200"))
import numpy as np
import gkrb200

bn = int(sys.argv[1]) if len(sys.argv) > 1 else 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = 1 << bn
peak = 6534.8
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
rng = np.random.default_rng(5)
def rnd(k):
    a = rng.integers(0, 1 << 63, size=(k, 4), dtype=np.uint64)
    a[:, 3] %= np.uint64(0x30644E72E131A029)
    return a
ctx = gkrb200.Context(0, bn)
tab, tab2, q, r = rnd(n), rnd(n), rnd(bn), rnd(1)[0]
ctx.set_profiling(True)
res = {}
def timed(name, cls, fn, alg_bytes):
    best = None
    for _ in range(reps):
        ctx.stats_reset()
        fn()
        ms = ctx.stats().kernel_ms[cls]
        best = ms if best is None or ms < best else best
    gbs = alg_bytes / (best * 1e-3) / 1e9
    res[name] = {"ms": best, "algorithmic_bytes": alg_bytes, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak}
    print("%-34s %8.3f ms  %7.1f MB algorithmic  %7.1f GB/s  = %.3f of %.1f GB/s" % (name, best, alg_bytes / 1e6, gbs, gbs / peak, peak), flush=True)
# K4 MultiLin.Fold (poly/multilin.go:26-36): read 2 x 32 B, write 32 B per output entry
timed("k_fold 2^%d -> 2^%d" % (bn, bn - 1), 3, lambda: gkrb200.poly.Fold(ctx, tab, r), 96 * (n // 2))
# K2 FoldedEqTable (poly/eq.go:41-59), one claim: one streaming write of the table (the sqrt(N)-size factor tables stay in L2)
timed("k_eq_small+k_eq_expand 2^%d (1 claim)" % bn, 1, lambda: gkrb200.poly.FoldedEqTable(ctx, q), 32 * n)
# layer 2's round kernel (identity gate, no fold): reads eq and X0 once
gate = gkrb200.gates.IdentityGate()
timed("k_round<identity> 2^%d (round 0)" % bn, 2, lambda: gkrb200.sumcheck.PartialEvals(ctx, tab, [tab2], gate), 64 * n)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "hbm_kernels_bn%d.json" % bn), "w"), indent=1)
ctx.close()
