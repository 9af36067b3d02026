"""Integer-pipe microbenchmarks (run under gpurun; optionally under ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gkr-mimc_b200"))
import gkrb200
ctx = gkrb200.Context(0, 10)
it = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
print("IMAD.WIDE.U32      : %.1f G MAC/s" % ctx.microbench(0, it)[0])
print("IMAD.WIDE.U32.X ch : %.1f G MAC/s" % ctx.microbench(2, it)[0])
print("fr_mul x2 full occ : %.1f G/s" % ctx.microbench(1, it)[0])
print("fr_mul x1 full occ : %.1f G/s" % ctx.microbench(3, it)[0])
for w in (4, 8, 12, 16, 24, 32):
    print("fr_mul x2 %2d warps/SM: %.1f G/s   x1: %.1f G/s" % (w, ctx.microbench(1 | (w << 8), it)[0], ctx.microbench(3 | (w << 8), it)[0]))
ctx.close()
