"""Per-role clock64() timeline of CTA 0 for one conv layer (needs MASIC_CONV_TRACE=1)."""
import os, sys
os.environ["MASIC_CONV_TRACE"] = "1"
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from masic_b200 import _lib
import conv_perf

names = sys.argv[1:] or ["g_a_conv2", "g_a_conv1(cin16)", "gmm_l0_fused"]
for name in names:
    plan = conv_perf.build(name, **conv_perf.LAYERS[name])
    for _ in range(2):
        plan.launch()
    torch.cuda.synchronize()
    buf = np.zeros((64, 16), dtype=np.int64)
    _lib.check(_lib.load().masic_conv_plan_trace(plan._h, buf.ctypes.data), "trace")
    t0 = buf[0, 0]
    print(f"== {name}: slots 0 mma.loop_top 1 mma.acc_empty_ok 2 mma.issue_done | 4 epi.top 5 epi.bias_ok 6 epi.acc_full 7 epi.pass1_done 8 epi.gdn_done 9 epi.end")
    print("  per-op stamps of tile 2 (delta to op start): wait_done fence_done issue_done syncwarp_done | start-to-start")
    prev = None
    for i in range(16):
        r = buf[32 + i]
        if r[0] == 0:
            break
        print(f"   op{i:2d}: {r[1]-r[0]:5d} {r[2]-r[0]:5d} {r[3]-r[0]:5d} {r[4]-r[0]:5d} | {0 if prev is None else r[0]-prev:6d}")
        prev = r[0]
    for it in range(min(3, 64)):
        if buf[it, 0] == 0:
            break
        r = buf[it] - t0
        print(f"tile {it}: mma top={r[0]:7d} go={r[1]:7d} issued={r[2]:7d} (issue {r[2]-r[1]:6d}) | epi top={r[4]:7d} bias={r[5]:7d} accfull={r[6]:7d} "
              f"p1={r[7]:7d} gdn={r[8]:7d} end={r[9]:7d} (epi {r[9]-r[6]:6d})")
