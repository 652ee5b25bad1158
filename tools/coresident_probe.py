"""Does leaving shared memory / registers free beside the persistent conv_tc CTA let the CUDA-core kernels of the other
lane / the other engines run under it?  Device-resident pairs/s of HSIC.pair_stream (the bench's `value` loop) for the
environment it is started in:   MASIC_CONV_SMEM_RESERVE=12288 python tools/coresident_probe.py [depth] [pairs]"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from masic_b200.hsic import HSIC  # noqa: E402

H, W = 1216, 2176
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 120
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = HSIC().eval().to(dev)
g = torch.Generator().manual_seed(1)
x1 = torch.rand(4, 3, H, W, generator=g).to(dev)
x2 = torch.rand(4, 3, H, W, generator=g).to(dev)
Hm = torch.eye(3)[None].repeat(4, 1, 1)
Hm[:, 0, 2] = torch.tensor([3.0, -5.0, 8.0, 1.5])
Hm = Hm.to(dev)
ps = net.pair_stream(H, W, dev, depth=depth)
for i in range(3 * depth):
    ps.submit(x1[i % 4:i % 4 + 1], x2[i % 4:i % 4 + 1], Hm[i % 4:i % 4 + 1], criterion=False)
ps.join()
best = 0.0
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        j = i % 4
        ps.submit(x1[j:j + 1], x2[j:j + 1], Hm[j:j + 1], criterion=False)
    ps.join()
    e1.record()
    torch.cuda.synchronize()
    best = max(best, n / (e0.elapsed_time(e1) / 1e3))
env = {k: v for k, v in os.environ.items() if k.startswith("MASIC_")}
print(f"depth {depth}: {best:.1f} pairs/s ({1e3 / best:.3f} ms per pair)  env={env}", flush=True)
