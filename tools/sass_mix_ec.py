#!/usr/bin/env python3
"""Static SASS opcode mix of the Groth16-side kernels (no GPU needed): cuobjdump -sass gkr-mimc_b200/libgkrb200ec.so -> profiles/r2_sass_ec_opcode_mix.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "gkr-mimc_b200", "libgkrb200ec.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
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
usage = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*(REG:\d+.*)", res):
    usage[m.group(1)] = m.group(2).strip()
names = dict(zip(funcs, subprocess.run(["c++filt"] + list(funcs), capture_output=True, text=True).stdout.splitlines()))
WANT = [("KAccum<ec::Curve<ec::FpBase", "G1 bucket accumulation: one mixed addition (madd-2008-s: 8 products + 2 squarings = 1 304 wide multiply-adds) per loop iteration"),
        ("KAccum<ec::Curve<ec::Fp2Base", "G2 bucket accumulation (Karatsuba Fp2 product = 3 base products, complex squaring = 2)"),
        ("KNtt<3, true>", "radix-8 DIF pass: 12 butterflies (12 products) per thread"), ("KNtt<3, false>", "radix-8 DIT pass"),
        ("KChunk<ec::Curve<ec::FpBase", "G1 window reduction (out-of-line multiplier)"), ("KScale", "coset scaling"), ("KCount", "digit histogram"),
        ("KScatter", "counting-sort scatter")]
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_ec_opcode_mix.txt")
with open(dst, "w") as f:
    f.write("# Static SASS opcode mix of the Groth16-side kernels (tools/sass_mix_ec.py: cuobjdump -sass gkr-mimc_b200/libgkrb200ec.so, sm_100a; no GPU needed).\n"
            "# These kernels have NOT been run on a GPU yet (DESIGN.md section 11): this is what the compiler produced, not a measurement.\n")
    for key, desc in WANT:
        hit = [k for k, v in names.items() if key in v]
        if not hit:
            f.write("== %s: not found\n" % key)
            continue
        ins = funcs[hit[0]]
        c = collections.Counter(re.sub(r"^@!?U?P\w+\s+", "", i).split()[0] for i in ins)
        wide = sum(v for o, v in c.items() if o.startswith("IMAD.WIDE"))
        junk = [(o, v) for o, v in c.most_common() if o.startswith("IMAD") and not o.startswith("IMAD.WIDE")]
        local = sum(v for o, v in c.items() if o.startswith(("LDL", "STL")))
        short = re.sub(r"void \(anonymous namespace\)::k_each<ec::", "", names[hit[0]]).split(",")[0]
        f.write("== %s\n   %s\n   %d instructions (%.1f KB); wide multiply-adds (IMAD.WIDE.U32[.X]) %d; other IMAD-family on the same pipe %d (%s); local-memory accesses %d\n   %s\n   %s\n" % (
            short, desc, len(ins), len(ins) * 16 / 1024, wide, sum(v for _, v in junk), ", ".join("%s %d" % x for x in junk) or "none", local,
            usage.get(hit[0], ""), ", ".join("%s %d" % x for x in c.most_common(14))))
print(open(dst).read())
