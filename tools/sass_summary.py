#!/usr/bin/env python
"""Instruction histogram per kernel of libcrnerf_b200.so (cuobjdump -sass): the mnemonics that show
which hardware paths a kernel uses - UTCHMMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTCBAR
(tcgen05.commit), UBLKCP (cp.async.bulk, either direction), UTMALDG/UTMASTG (tensor-map TMA, not used: every bulk copy
here is a 1-D pre-swizzled image), SYNCS (mbarrier), LDG/STG widths, USETMAXREG.

  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cr-nerf-pytorch_b200", "crnerf_b200", "libcrnerf_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "USETMAXREG",
        "LDG.256", "LDG.E.128", "LDG", "STG.256", "STG.E.128", "STG", "LDS", "STS", "LDL", "STL", "SHFL", "F2FP", "MUFU", "FFMA",
        "HFMA2", "BAR"]
kern, hist, total = None, {}, {}
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").replace("void ", "")
        kern = re.sub(r"\(.*", "", kern)
        while kern in hist:
            kern += "'"
        hist[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        if op.startswith(("LDG", "STG")) and ".256" in op:   # e.g. LDG.E.NA.ENL2.256.CONSTANT, STG.E.EF.ENL2.256
            hist[kern][op[:3] + ".256"] += 1
            continue
        for k in KEYS:
            if op.startswith(k):
                hist[kern][k] += 1
                break
print(f"# {os.path.basename(lib)}: SASS instruction counts per kernel (static), sm_100a")
print("# kernel | total | " + " ".join(KEYS))
for k in sorted(hist, key=lambda n: -total[n]):
    if total[k] < 40:
        continue
    cells = " ".join(f"{key}={hist[k][key]}" for key in KEYS if hist[k][key])
    print(f"{k} | {total[k]} | {cells}")
