"""SASS mnemonic counts per kernel of the built library -> profiles/rN_sass_summary.txt

    python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "masic_b200" / "libmasic_b200.so"
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "USETMAXREG", "FFMA2", "FMUL2",
        "FADD2", "MUFU.RSQ", "F2FP", "DFMA"]
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        c = funcs[cur]
        base = op.split(".")[0]
        if base == "UTCHMMA":
            c["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
        elif op.startswith("MUFU.RSQ"):
            c["MUFU.RSQ"] += 1
        elif base in KEYS:
            c[base] += 1
names = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
print("# SASS mnemonic counts per kernel of masic_b200/libmasic_b200.so (cuobjdump -sass; sm_100a cubins only)")
print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store,")
print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, USETMAXREG = setmaxnreg, FFMA2/FMUL2/FADD2 = packed fp32 pairs")
print("# (kernels without any of the listed mnemonics are omitted)\n")
for (mangled, c), name in zip(funcs.items(), names):
    items = [f"{k}={c[k]}" for k in KEYS if c[k]]
    if items:
        print(name)
        print("    " + "  ".join(items))
