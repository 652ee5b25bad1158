"""Time the CQE engine (Independent_EN) on a 1216x2176 pair: per-step CUDA-event profile + graph replay.
    python tools/cqe_perf.py [H W]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from masic_b200.cqe import Independent_EN  # noqa: E402

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1216, 2176)
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = Independent_EN().eval().to(dev)
eng = net.engine_for(1, H, W, dev)
g = torch.Generator().manual_seed(1)
eng.x1.copy_(torch.rand(1, 3, H, W, generator=g))
eng.x2.copy_(torch.rand(1, 3, H, W, generator=g))
eng.Hm.copy_(torch.tensor([[1.0, 0.01, 20.0], [0.0, 1.0, 3.0], [1e-6, 0.0, 1.0]]))
prof = eng.profile_steps(3)
tot = sum(t for _, t in prof)
for n, t in prof:
    pl = eng.plans.get(n)
    extra = f"  {pl.flops / t / 1e9:7.1f} TF/s  work={pl.work_items}" if pl is not None else ""
    print(f"{t:8.4f} ms {100 * t / tot:5.1f}%  {n}{extra}")
conv_ms = sum(t for n, t in prof if n in eng.plans)
print(f"total {tot:.3f} ms; conv {conv_ms:.3f} ms; plan flops {eng.flops / 1e9:.1f} GFLOP -> {eng.flops / conv_ms / 1e9:.1f} TF/s in convs")
eng.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    eng.run()
e1.record()
torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1) / 5:.3f} ms per pair")
