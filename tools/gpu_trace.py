"""Stage timestamps (SM clock) of the small-round kernel; needs a library built with -DGKR_TRACE (make TRACE=1)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gkr-mimc_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import gkrb200
from gkrb200._lib import lib
ctx = gkrb200.Context(0, 16)
rng = np.random.default_rng(0)
ark = gkrb200.common.SetUint64([5])[0]
for bn in (1, 4, 5, 8, 12):
    n = 1 << bn
    L = gkrb200.common.RandomFrArray(n); R = L[::-1].copy(); q = gkrb200.common.RandomFrArray(bn).reshape(1, bn, 4)
    for rep in range(2):
        gkrb200.sumcheck.Prove(ctx, [L, R], q, None, gkrb200.gates.CipherGate(ark))
    t = (ctypes.c_longlong * 16)()
    lib().gkrb200_trace_get(t)
    t = list(t)
    names = ["zero smem", "loads+fold+shfl", "compute", "(acc)+sync", "block tree", "grid stage", "publish"]
    print("bn=%d last round: " % bn + ", ".join("%s %d cyc" % (nm, t[i + 1] - t[i]) for i, nm in enumerate(names)) + " | total %d cyc = %.2f us" % (t[7] - t[0], (t[7] - t[0]) / 1965.0))
ctx.close()
