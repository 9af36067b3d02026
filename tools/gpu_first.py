"""First GPU contact: integer-pipe microbenchmarks + a few timed proofs (not a bench; a sanity probe)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gkr-mimc_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import gkrb200

max_bn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = gkrb200.Context(0, max_bn)
for kind, name, unit in ((0, "IMAD.WIDE.U32", "G wide-MAC/s"), (1, "fr_mul chains", "G Fr-mul/s")):
    for iters in (200, 2000):
        rate, ms = ctx.microbench(kind, iters)
        print("microbench %-14s iters=%5d  %.1f %s  (%.3f ms)" % (name, iters, rate, unit, ms), flush=True)
c = gkrb200.MimcCircuit(ctx)
for bn in (10, 16, max_bn):
    n = 1 << bn
    rng = np.random.default_rng(bn)
    key = gkrb200.common.RandomFrArray(n); msg = key[::-1].copy(); q = gkrb200.common.RandomFrArray(bn)
    for rep in range(2):
        ctx.stats_reset()
        ctx.set_profiling(rep == 1)
        t0 = time.time(); a = c.Assign(key, msg); t1 = time.time(); p = gkrb200.gkr.Prove(c, a, q); t2 = time.time()
        s = ctx.stats()
        print("bn=%d rep=%d assign %.1f ms prove %.1f ms  -> %.0f hashes/s | transcript %.1f ms wait %.1f ms launches %d rounds %d kernel_ms %s" % (
            bn, rep, (t1-t0)*1e3, (t2-t1)*1e3, n/(t2-t0), s.transcript_ms, s.wait_ms, s.launches_total, s.rounds,
            ["%.2f" % x for x in s.kernel_ms[:6]]), flush=True)
        if rep == 1 and s.kernel_ms[2] > 0:
            print("   round kernels: %.1f G Fr-mul/s, %.1f GB/s ; assign: %.1f G Fr-mul/s" % (
                s.fr_mul_round / s.kernel_ms[2] / 1e6, s.bytes_round / s.kernel_ms[2] / 1e6, s.fr_mul_assign / max(s.kernel_ms[0], 1e-9) / 1e6), flush=True)
ctx.close()
