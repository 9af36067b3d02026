"""One Assign + gkr.Prove at 2^bn (for ncu captures)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gkr-mimc_b200"))
import gkrb200
bn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = gkrb200.Context(0, bn)
c = gkrb200.MimcCircuit(ctx)
key = gkrb200.common.RandomFrArray(1 << bn); msg = key[::-1].copy(); q = gkrb200.common.RandomFrArray(bn)
a = c.Assign(key, msg)
t = time.time(); p = gkrb200.gkr.Prove(c, a, q); print("prove %.1f ms" % ((time.time() - t) * 1e3))
ctx.close()
