"""Decompose the engine-vs-oracle PSNR deviation of the left view in the trained regime: flips vs decoder arithmetic.
Trains a model with the repo's training step, then per image: PSNR of the oracle, of the engine, and of the ORACLE's
decoder fed with the ENGINE's symbols (same symbols => the difference to the engine is decoder arithmetic only).
(Test infrastructure: uses oracle/.)   python tools/trained_diag.py [steps]"""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from masic_b200.hsic import HSIC  # noqa: E402
from oracle import hsic as OH  # noqa: E402
from tools.train_regime import smooth_pairs, train_to_psnr  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = HSIC().to(dev)
steps, ps = train_to_psnr(net, dev, target_db=27.5, max_steps=int(sys.argv[1]) if len(sys.argv) > 1 else 750, size=(256, 256),
                          lr=3e-4, lmbda=0.05)
print("trained", steps, ps)
net.eval()
oracle = OH.OracleHSIC(128, 192, 5).eval()
oracle.load_state_dict({k: v.detach().cpu() for k, v in net.state_dict().items()})
oracle = oracle.to(dev)                       # fp32 torch modules on the GPU (cuDNN, TF32 off by default)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
psnr = lambda a, b: 10 * math.log10(1.0 / float(torch.mean((a.double() - b.double()) ** 2)))  # noqa: E731
for (h, w, n) in ((256, 384, 8), (512, 512, 4), (1216, 2176, 2)):
    acc = [0.0, 0.0, 0.0]
    for i in range(n):
        g = torch.Generator().manual_seed(9 + i)
        x1, x2, Hm = smooth_pairs(1, h, w, g)
        x1, x2, Hm = x1.to(dev), x2.to(dev), Hm.to(dev)
        with torch.no_grad():
            ref = oracle(x1, x2, Hm)
            out = net(x1, x2, Hm)
            mixed = OH._run_decoder(oracle.decoder1, out["y1_hat"])
        p = (psnr(ref["x1_hat"], x1), psnr(out["x1_hat"], x1), psnr(mixed, x1))
        flips = int((out["y1_hat"] != ref["y1_hat"]).sum())
        print(f"{h}x{w} #{i}: oracle {p[0]:.4f}  engine {p[1]:.4f} ({p[1] - p[0]:+.4f})  oracle decoder on engine symbols {p[2]:.4f} "
              f"({p[2] - p[0]:+.4f}); engine - same-symbol oracle {p[1] - p[2]:+.5f}; flips {flips}; "
              f"rms(engine - mixed) {float((out['x1_hat'] - mixed).pow(2).mean().sqrt()):.2e}")
        for k in range(3):
            acc[k] += p[k] / n
    print(f"{h}x{w} mean: engine - oracle {acc[1] - acc[0]:+.5f}; flips only {acc[2] - acc[0]:+.5f}; decoder arithmetic {acc[1] - acc[2]:+.5f}")
