"""Fixed cost of a conv_tc launch: one-tile plans timed alone, back to back, and alternating with a small elementwise
kernel (does the 227 KB shared-memory carve-out switch cost anything?).  python tools/launch_overhead.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from masic_b200.convplan import ConvPlan  # noqa: E402

dev = torch.device("cuda:0")


def plan(h, w, cin, cout, n_tile):
    x = torch.randn(1, h, w, cin, device=dev).to(torch.float16)
    wt = torch.randn(cout, cin, 1, 1, device=dev) * 0.05
    out = torch.empty(1, h, w, cout, device=dev, dtype=torch.float16)
    return ConvPlan(ksize=1, x=x, c_in=cin, weight=wt, bias=torch.zeros(cout, device=dev), c_out=cout, n_tile=n_tile, out=out)


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


small = torch.zeros(1 << 16, device=dev)
for name, p in (("1 tile K=64 N=16", plan(16, 8, 64, 16, 16)), ("148 tiles K=64 N=128", plan(16 * 148, 8, 64, 128, 128)),
                ("148 tiles K=768 N=192", plan(16 * 148, 8, 768, 192, 192)),
                ("1530 tiles K=768 N=192 (gmm l0 shape)", plan(76, 136, 768, 3456, 192))):
    a = timeit(p.launch)
    b = timeit(lambda: (p.launch(), small.add_(1.0)))
    c = timeit(lambda: small.add_(1.0))
    print(f"{name:40s} back-to-back {a:7.2f} us/launch; alternating with a small kernel {b:7.2f} us/pair (small alone {c:.2f} us)")

# the latency-bound small layers of the step, warm and back to back inside one graph (tools/conv_perf.py shapes)
import importlib.util
spec = importlib.util.spec_from_file_location("conv_perf", str(Path(__file__).resolve().parent / "conv_perf.py"))
cp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(cp)
for name in ("mb_ha2_n32", "mb_ha3_n32", "mb_ga4_n192", "mb_ha1_n128", "mb_hs2_n192", "mb_hs1_n128"):
    p = cp.build(name, **cp.MICRO[name])
    print(f"{name:40s} back-to-back {timeit(p.launch):7.2f} us/launch  work={p.work_items}")
