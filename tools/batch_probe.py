"""Does a batch-2 engine amortise the latency-bound launches?  Graph replay of one engine at batch 1 / 2 / 3 (1216x2176),
ms per pair, and two / three such engines on their own streams.   python tools/batch_probe.py"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from masic_b200.engine import HSICEngine  # noqa: E402
from masic_b200.hsic import HSIC  # noqa: E402

H, W = 1216, 2176
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = HSIC().eval().to(dev)
net.engine_for(1, 64, 64, dev)          # applies the MaskedConv2d weight masking
for B in (1, 2, 3):
    for n_eng in (1, 2, 3):
        if B * n_eng > 6:
            continue
        engs = [HSICEngine(net.state_dict(), B, H, W, dev, net.N, net.M, net.K, use_graph=True) for _ in range(n_eng)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_eng)]
        g = torch.Generator().manual_seed(1)
        for e in engs:
            e.x1.copy_(torch.rand(B, 3, H, W, generator=g)); e.x2.copy_(torch.rand(B, 3, H, W, generator=g))
            e.Hm.copy_(torch.eye(3)[None].repeat(B, 1, 1))
        for _ in range(3):
            for e, s in zip(engs, streams):
                with torch.cuda.stream(s):
                    e.run()
        torch.cuda.synchronize()
        n = 12
        t0 = time.perf_counter()
        for _ in range(n):
            for e, s in zip(engs, streams):
                with torch.cuda.stream(s):
                    e.run()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / (n * n_eng * B)
        print(f"batch {B} x {n_eng} engine(s): {dt * 1e3:.3f} ms per pair  ({1 / dt:.1f} pairs/s)", flush=True)
        del engs
        torch.cuda.empty_cache()
