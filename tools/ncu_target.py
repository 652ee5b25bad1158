"""Short ncu target: one warm eager HSIC step at 1216x2176, then the selected steps once each
between cudaProfilerStart/Stop (run under `ncu --profile-from-start off --set full ...`).

    python tools/ncu_target.py [substring ...]      # default: the heaviest launch of every kernel family
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from masic_b200.hsic import HSIC  # noqa: E402

DEFAULT = ["L.g_a.conv1+gdn", "L.g_a.conv2+gdn", "L.g_a.conv3+gdn", "L.g_a.conv4", "L.g_s.deconv3+igdn",
           "L.g_s.deconv4(col2im)", "R.g_s.deconv4(col2im)", "L.gmm.l0", "L.gmm.l1", "L.gmm.l2(sigma|means)",
           "L.gmm_likelihood", "R.warp(x1)", "R.pre_conv+pre_gdn", "R.after_conv", "L.h_a.conv2", "L.h_s.conv3x3",
           "L.context", "L.entropy_bottleneck", "x1.pack_nhwc", "L.latent_prep"]

want = sys.argv[1:] or DEFAULT
if want == ["ALL"]:          # every launch of one step (e.g. `--metrics dram__bytes_read.sum,dram__bytes_write.sum`)
    want = [""]
torch.manual_seed(0)
net = HSIC().eval().cuda()
eng = net.engine_for(1, 1216, 2176, torch.device("cuda:0"))
g = torch.Generator().manual_seed(100)
eng.x1.copy_(torch.rand(1, 3, 1216, 2176, generator=g))
eng.x2.copy_(torch.rand(1, 3, 1216, 2176, generator=g))
for _ in range(2):
    eng._launch_all(concurrent=False)
torch.cuda.synchronize()
sel = [(n, fn) for n, fn in eng.steps if any(w in n for w in want)]
torch.cuda.profiler.start()
for n, fn in sel:
    fn()
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled steps:", [n for n, _ in sel])
