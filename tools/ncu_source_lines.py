"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line:
    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_source_lines.py src.csv [kernel_index] [top_n]
Prints, per source line, stall samples and executed warp instructions (top N by samples)."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
# split per kernel: each starts with a "File Path" row (cuda view) — collect blocks whose header row starts with "Line No"
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Line No":
        cur = {"hdr": r, "rows": []}
        blocks.append(cur)
    elif cur is not None and r and r[0] not in ("File Path", "Function Name", "Kernel Name"):
        cur["rows"].append(r)
print(f"{len(blocks)} source blocks")
b = blocks[kidx]
h = b["hdr"]
iS, iI = h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
agg = defaultdict(lambda: [0, 0, "", defaultdict(int)])
line, src = None, ""
for r in b["rows"]:
    if r[0] != "":
        line, src = r[0], r[1]
    try:
        s, n = int(r[iS] or 0), int(r[iI] or 0)
    except ValueError:
        continue
    if r[2] in ("", "-"):      # the cuda-line summary row itself (already the sum of its sass rows) — skip to avoid double count
        continue
    a = agg[line]
    a[0] += s; a[1] += n; a[2] = src
    for i, nme in stall_cols:
        try:
            a[3][nme] += int(r[i] or 0)
        except ValueError:
            pass
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
print(f"total samples {tot_s}, warp instructions {tot_i}")
for line, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    top = sorted(a[3].items(), key=lambda kv: -kv[1])[:3]
    print(f"L{line:>5} samp {a[0]:6d} ({100 * a[0] / max(tot_s, 1):4.1f}%) inst {a[1]:9d} ({100 * a[1] / max(tot_i, 1):4.1f}%)  "
          f"{' '.join(f'{k[6:]}={v}' for k, v in top):40s} | {a[2].strip()[:90]}")
