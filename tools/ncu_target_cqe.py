"""Short ncu target for the CQE engine: one warm eager step at 1216x2176, then the selected steps once each between
cudaProfilerStart/Stop (run under `ncu --profile-from-start off --set full ...`).
    python tools/ncu_target_cqe.py [substring ...]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from masic_b200.cqe import Independent_EN  # noqa: E402

DEFAULT = ["mask2weights_L", "L.blend_images", "L.conv1", "L.EB1.RB1.conv1", "L.EB1.RB1.conv2+skip", "L.feature_fuse",
           "L.EB2.RB1.conv1", "L.EB2.RB3.conv2+skips", "L.EB3.RB1.conv1", "L.EB3.RB1.conv2+skip", "L.EB3.RB3.conv2+skips",
           "L.conv2", "L.residual_image"]
want = sys.argv[1:] or DEFAULT
torch.manual_seed(0)
net = Independent_EN().eval().cuda()
eng = net.engine_for(1, 1216, 2176, torch.device("cuda:0"))
g = torch.Generator().manual_seed(1)
eng.x1.copy_(torch.rand(1, 3, 1216, 2176, generator=g))
eng.x2.copy_(torch.rand(1, 3, 1216, 2176, generator=g))
eng.Hm.copy_(torch.tensor([[1.0, 0.01, 20.0], [0.0, 1.0, 3.0], [1e-6, 0.0, 1.0]]))
for _ in range(2):
    eng._launch_all(concurrent=False)
torch.cuda.synchronize()
sel = [(n, fn) for n, fn in eng.steps if any(n == w or (w.endswith("_L") and n.startswith(w)) for w in want)]
torch.cuda.profiler.start()
for n, fn in sel:
    fn()
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled steps:", [n for n, _ in sel])
