"""Experiment: throughput of N independent batch-1 engines replayed concurrently on N streams of one GPU.
    python tools/dual_engine.py [n_engines]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from masic_b200.engine import HSICEngine  # noqa: E402
from masic_b200.hsic import HSIC  # noqa: E402

n_eng = int(sys.argv[1]) if len(sys.argv) > 1 else 2
H, W = 1216, 2176
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = HSIC().eval().to(dev)
with torch.no_grad():
    for cp in (net.context_prediction1, net.context_prediction2):
        cp.weight.data *= cp.mask
sd = net.state_dict()
engs = [HSICEngine(sd, 1, H, W, dev) for _ in range(n_eng)]
streams = [torch.cuda.Stream() for _ in range(n_eng)]
g = torch.Generator().manual_seed(1)
for e in engs:
    e.x1.copy_(torch.rand(1, 3, H, W, generator=g)); e.x2.copy_(torch.rand(1, 3, H, W, generator=g))
    e.Hm.copy_(torch.tensor([[1.0, 0.01, 20.0], [0.0, 1.0, 3.0], [1e-6, 0.0, 1.0]]))
for e, s in zip(engs, streams):
    with torch.cuda.stream(s):
        e.run(); e.run()
torch.cuda.synchronize()
steps = 40
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in streams:
    s.wait_stream(torch.cuda.current_stream())
for i in range(steps):
    k = i % n_eng
    with torch.cuda.stream(streams[k]):
        engs[k].run()
for s in streams:
    torch.cuda.current_stream().wait_stream(s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"{n_eng} engine(s): {ms:.3f} ms per pair = {1e3 / ms:.1f} pairs/s")
