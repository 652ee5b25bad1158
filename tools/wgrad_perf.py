"""Warm CUDA-event time of the tensor-core weight-gradient plan on the training step's layer shapes (512x896, batch 2).
    python tools/wgrad_perf.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from masic_b200.convplan import WgradPlan  # noqa: E402

dev = torch.device("cuda:0")
B, H, W = 2, 512, 896
shapes = [  # name, ksize, stride, (h_lo, w_lo), c_lo, c_hi
    ("conv2 5x5 s2 128<-128 @128x224", 5, 2, (H // 4, W // 4), 128, 128),
    ("conv3 5x5 s2 128<-128 @64x112", 5, 2, (H // 8, W // 8), 128, 128),
    ("conv4 5x5 s2 192<-128 @32x56", 5, 2, (H // 16, W // 16), 192, 128),
    ("gdn 1x1 128<-128 @256x448", 1, 1, (H // 2, W // 2), 128, 128),
    ("gmm.l0 1x1 3456<-768 @32x56", 1, 1, (H // 16, W // 16), 3456, 768),
    ("ctx 5x5 s1 384<-192 @32x56", 5, 1, (H // 16, W // 16), 384, 192),
]
for name, k, s, (hl, wl), cl, ch in shapes:
    lo = torch.randn(B, hl, wl, cl, device=dev).to(torch.bfloat16)
    hi = torch.randn(B, hl * s, wl * s, ch, device=dev).to(torch.bfloat16)
    dw = torch.zeros(cl, ch, k, k, device=dev)
    pl = WgradPlan(ksize=k, stride=s, lo=lo, c_lo=cl, hi=hi, c_hi=ch, dw=dw)
    for _ in range(3):
        pl.launch()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pl.launch(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name:38s} {ms * 1e3:8.1f} us  {pl.flops / ms / 1e9:7.1f} TF/s  ctas={pl.n_ctas}", flush=True)
