"""Where the end-to-end figure loses against the device-resident one: PairStream at 1216x2176 with device / pinned-host
inputs, with and without the criterion read-back.  python tools/e2e_diag.py"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from masic_b200.hsic import HSIC  # noqa: E402

H, W, N = 1216, 2176, 40
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = HSIC().eval().to(dev)
g = torch.Generator().manual_seed(1)
xh = [torch.rand(1, 3, H, W, generator=g).pin_memory() for _ in range(8)]
xu = [(t * 255).round().to(torch.uint8).pin_memory() for t in xh]
xd = [t.to(dev) for t in xh]
hh = torch.eye(3)[None].pin_memory()
hd = hh.to(dev)
ps = net.pair_stream(H, W, dev, depth=3)


def run(src, h, criterion, lag=2):
    pend = []
    for i in range(N):
        pend.append(ps.submit(src[(2 * i) % 8], src[(2 * i + 1) % 8], h, criterion=criterion))
        if len(pend) > lag:
            ps.result(pend.pop(0))
    for t in pend:
        ps.result(t)


for name, src, h, crit in (("device inputs, no criterion", xd, hd, False), ("device inputs, criterion + read-back", xd, hd, True),
                           ("host fp32 inputs, no criterion", xh, hh, False), ("host fp32 inputs, criterion + read-back", xh, hh, True),
                           ("host uint8 inputs, no criterion", xu, hh, False), ("host uint8 inputs, criterion + read-back", xu, hh, True)):
    run(src, h, crit)
    torch.cuda.synchronize()
    reps = []
    for _ in range(6):
        t0 = time.perf_counter()
        run(src, h, crit)
        torch.cuda.synchronize()
        reps.append((time.perf_counter() - t0) / N)
    best = min(reps)
    print(f"{name:42s} {best * 1e3:.3f} ms/pair  {1 / best:.1f} pairs/s   all: " + " ".join(f"{1 / r:.0f}" for r in reps))
