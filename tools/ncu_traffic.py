"""ncu per-launch DRAM traffic csv -> profiles/rN_conv_traffic.json (read by bench.py for roofline.traffic).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \\
        --profile-from-start off --csv --log-file gpurun_out/traffic.csv python tools/ncu_target.py ALL
    python tools/ncu_traffic.py gpurun_out/traffic.csv > profiles/r1_conv_traffic.json
"""
import csv
import json
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
i_id, i_k, i_m, i_v = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
per = defaultdict(dict)
for r in rows[1:]:
    per[(r[i_id], r[i_k])][r[i_m]] = float(r[i_v].replace(",", ""))
conv = [v for (_, k), v in per.items() if "conv_tc_kernel" in k or "deconv_img_kernel" in k]     # the tcgen05 kernels
allk = list(per.values())
b = lambda v: v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0)
out = {
    "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, "
              "python tools/ncu_target.py ALL (one eager 1216x2176 step, every launch)",
    "conv_tc_launches": len(conv),
    "conv_tc_dram_bytes_per_step": sum(b(v) for v in conv),
    "conv_tc_dram_bytes_per_launch": sum(b(v) for v in conv) / max(len(conv), 1),
    "conv_tc_us_per_step_under_ncu": sum(v.get("gpu__time_duration.sum", 0.0) for v in conv) / 1e3,
    "all_launches": len(allk),
    "all_dram_bytes_per_step": sum(b(v) for v in allk),
    "all_us_per_step_under_ncu": sum(v.get("gpu__time_duration.sum", 0.0) for v in allk) / 1e3,
}
print(json.dumps(out, indent=1))
