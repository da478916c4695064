"""SASS evidence per kernel of the built library (no GPU needed): counts of the Blackwell-specific
mnemonics north_star / SURVEY 8(d) name -- UBLKCP (1-D TMA bulk copy), FFMA2/FADD2/FMUL2 (packed FP32),
FMNMX3, SYNCS (mbarrier), MUFU, and for the exponential window's tensor-core pass UTCHMMA (tcgen05.mma),
UTCBAR (tcgen05.commit), LDTM (tcgen05.ld), UTCATOMSWS (tcgen05.alloc), LDGSTS (cp.async) -- plus a few lines of
each around the first occurrence.
usage: python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pyfstat_b200", "libtcw_b200.so")
KEYS = ("UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FMNMX3", "MUFU.RCP", "MUFU.EX2", "DFMA", "DADD", "F2F",
        "LDS.128", "LDS.64", "STG.E.128", "STG.E", "LDG.E.EF.128", "ATOMS", "ATOMG", "REDG", "BAR.SYNC",
        "UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "LDGSTS")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
print(f"cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo)")
print("counts of selected SASS mnemonics per kernel (static instruction counts, not executions)\n")
kernels = re.split(r"\n\s*Function : ", out)[1:]
for k in kernels:
    name = k.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    lines = [ln for ln in k.split("\n") if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln)]
    ops = collections.Counter()
    for ln in lines:
        body = ln.split("*/", 1)[1]
        for key in KEYS:
            if re.search(r"\b" + re.escape(key) + r"\b", body) or (key in body and "." in key):
                ops[key] += 1
    if "peak_kernel" in dem:
        continue
    print(f"== {dem}: {len(lines)} instructions")
    print("   " + "  ".join(f"{key}={ops[key]}" for key in KEYS if ops[key]))
    for key in ("UBLKCP", "FFMA2", "SYNCS", "MUFU.EX2", "UTCHMMA", "LDTM", "LDGSTS"):
        for ln in lines:
            if key in ln:
                print("     e.g. " + ln.split("*/", 1)[1].split(";")[0].strip())
                break
    print()
