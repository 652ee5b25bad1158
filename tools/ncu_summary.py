"""Summarise ncu outputs into small text files for profiles/.

    python tools/ncu_summary.py launches gpurun_out/r1b_launches.csv   > profiles/r1_launches.md
    python tools/ncu_summary.py full     gpurun_out/r1b_full.ncu-rep   > profiles/r1_ncu_full.md
"""
import csv
import subprocess
import sys
from collections import OrderedDict, defaultdict

KEYS = OrderedDict([
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "MB"),
    ("dram__bytes_write.sum", "MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor % (active)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor % (elapsed)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__cycles_elapsed.max", "cycles"),
])


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = defaultdict(lambda: [0, 0.0])
    order = []
    for r in rows[1:]:
        name = r[ik].split("(")[0].replace("void ", "")
        if "spin_kernel" in name:           # torch.cuda._sleep in front of the per-kernel attribution pass: not work
            continue
        if name.startswith("at::") or "elementwise" in name:
            name = "torch: " + name[:60]
        ns = float(r[iv].replace(",", ""))
        agg[name][0] += 1
        agg[name][1] += ns
        order.append((name, r[ig], ns))
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list ({len(order)} launches captured, serialised & cold-cache: compare SHARES)\n")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {ns / 1e3:.1f} | {100 * ns / tot:.1f}% |")
    print(f"\ntotal {tot / 1e3:.1f} us\n\n## every launch in order (us)\n")
    for i, (name, grid, ns) in enumerate(order):
        print(f"{i:4d} {ns / 1e3:9.2f}  {grid:>16s}  {name}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full, one launch per kernel family (1216x2176 pair, eager step, warm)\n")
    cols = [k for k in KEYS if k in idx]
    print("| # | kernel | grid | " + " | ".join(KEYS[k] for k in cols) + " |")
    print("|---|---|---|" + "---:|" * len(cols))
    for n, r in enumerate(rows[2:]):
        vals = []
        for k in cols:
            v, u = r[idx[k]], units[idx[k]]
            try:
                f = float(v.replace(",", ""))
                if k.startswith("dram__bytes") or "xbar2l1tex" in k:
                    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                    v = f"{f * scale:.1f}"
                elif k == "gpu__time_duration.sum":
                    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
                    v = f"{f * scale:.1f}"
                else:
                    v = f"{f:.1f}" if f != int(f) else str(int(f))
            except ValueError:
                pass
            vals.append(v)
        print(f"| {n} | `{r[idx['Kernel Name']].split('(')[0]}` | {r[idx['Grid Size']] if 'Grid Size' in idx else ''} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
