#!/usr/bin/env python3
"""Static SASS opcode mix of the shipped kernels (no GPU needed): cuobjdump -sass gkr-mimc_b200/libgkrb200.so -> profiles/r2_sass_opcode_mix.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "gkr-mimc_b200", "libgkrb200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append(m.group(1))
names = dict(zip(funcs, subprocess.run(["c++filt"] + list(funcs), capture_output=True, text=True).stdout.splitlines()))
WANT = [("k_round_cf<true, 7, 1, 256, 2, true, true>", "shipped fold round (rounds >= 1, one thread per pair, inlined multiplier, constant-multiplier fold)"),
        ("k_round_cf<false, 7, 1, 256, 2, true, false>", "shipped round 0 (no fold)"),
        ("k_round_cf<true, 7, 1, 128, 3, false, false>", "out-of-line multiplier build of the fold round (round-1 design; A/B reference via GKRB200_OPT_INLINE_MIN_PAIRS)"),
        ("k_round_cf<true, 7, 8, 128, 4, false, false>", "8 lanes per pair (rounds of <= 8192 pairs)"),
        ("gkr::k_mimc_assign(", "K1"), ("gkr::k_fold(", "K4 standalone")]
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_opcode_mix.txt")
with open(dst, "w") as f:
    f.write("# Static SASS opcode mix of the shipped kernels (tools/sass_mix.py: cuobjdump -sass gkr-mimc_b200/libgkrb200.so, sm_100a; no GPU needed).\n"
            "# The body of a one-thread-per-pair round kernel runs once per pair, so its static counts are (up to prologue/epilogue) the per-pair\n"
            "# dynamic counts the ncu source pages report (profiles/r1_ncu_k_round_cf_fold_opcode_mix.txt for round 1's build).\n")
    for key, desc in WANT:
        hit = [k for k, v in names.items() if key in v]
        if not hit:
            f.write("== %s: not found\n" % key)
            continue
        ins = funcs[hit[0]]
        c = collections.Counter(re.sub(r"^@!?U?P\w+\s+", "", i).split()[0] for i in ins)
        wide = sum(v for o, v in c.items() if o.startswith("IMAD.WIDE"))
        junk = [(o, v) for o, v in c.most_common() if o.startswith("IMAD") and not o.startswith("IMAD.WIDE")]
        f.write("== %s\n   %s\n   %d instructions (%.1f KB); wide multiply-adds (IMAD.WIDE.U32[.X]) %d; other IMAD-family on the same pipe %d (%s)\n   %s\n" % (
            names[hit[0]].split("(")[0].replace("void gkr::", "").replace("gkr::", ""), desc, len(ins), len(ins) * 16 / 1024, wide, sum(v for _, v in junk),
            ", ".join("%s %d" % x for x in junk) or "none", ", ".join("%s %d" % x for x in c.most_common(14))))
print(open(dst).read())
