"""Per-step (per kernel launch) CUDA-event profile of one HSIC forward at 1216x2176."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from masic_b200.hsic import HSIC

h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1216, 2176)
torch.manual_seed(0)
net = HSIC().eval().cuda()
eng = net.engine_for(1, h, w, torch.device("cuda:0"))
prof = eng.profile_steps(iters=5)
tot = sum(ms for _, ms in prof)
for n, ms in prof:
    pl = eng.plans.get(n)
    extra = f"  {pl.flops / ms / 1e9:7.1f} TF/s  work={pl.work_items}" if pl else ""
    print(f"{ms:8.4f} ms {100 * ms / tot:5.1f}%  {n}{extra}")
print(f"total {tot:.3f} ms; conv {sum(ms for n, ms in prof if n in eng.plans):.3f} ms")
