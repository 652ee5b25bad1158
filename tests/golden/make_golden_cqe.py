"""Golden fixture for the cross quality enhancement network, generated from the UNMODIFIED reference
(`MASIC.Independent_EN`, coremasic/mywork/MASIC.py:1436-1501).  Run where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden_cqe.py

Pins oracle/cqe.py (tests/test_oracle_cqe_pinned.py) and, through it, the CUDA path (tests/test_cqe_gpu.py).
kornia is oracle/shims/kornia (restated 0.5.0): the warp itself stays unpinned by the reference.
"""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
from oracle import refimport  # noqa: E402
from oracle.hsic import synthetic_homography  # noqa: E402

torch.set_num_threads(8)


def sha(t):
    return hashlib.sha256(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()


def main():
    MASIC = refimport.import_masic()
    torch.manual_seed(0)
    net = MASIC.Independent_EN().eval()
    sd = net.state_dict()
    layout = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()]
    (HERE / "cqe_layout.json").write_text(json.dumps(
        {"entries": layout, "params": sum(p.numel() for p in net.parameters()),
         "param_sha256": {k: sha(sd[k]) for k in ("EBl1.RB1.conv1.weight", "EBr3.RB3.conv2.bias", "conv0.weight",
                                                  "conv1.bias", "conv2.weight", "mask2weights_unit.maskconv.6.weight")}},
        indent=0))
    for tag, h, w, gain in (("init_64x96", 64, 96, 1.0), ("gain1p3_96x128", 96, 128, 1.3)):
        torch.manual_seed(0)
        net = MASIC.Independent_EN().eval()
        if gain != 1.0:           # random init gives a near-zero correction: scale the convs so every stage matters
            with torch.no_grad():
                for n_, p_ in net.named_parameters():
                    if n_.endswith("weight") and p_.dim() == 4 and not n_.startswith("mask2weights"):
                        p_.mul_(gain)
        g = torch.Generator().manual_seed(200)
        x1 = torch.rand(2, 3, h, w, generator=g)
        x2 = torch.rand(2, 3, h, w, generator=g)
        Hm = synthetic_homography(2, seed=5)
        Hm[:, 0, 2] *= 0.25        # keep the overlap large on these small images
        with torch.no_grad():
            o = net(x1, x2, Hm)
        cy, cx = h // 2 - 24, w // 2 - 24
        np.savez_compressed(HERE / f"cqe_forward_{tag}.npz", h=h, w=w, gain=gain, x_seed=200, h_seed=5, H=Hm.numpy(),
                            x1_hat_crop=o["x1_hat"][:, :, cy:cy + 48, cx:cx + 48].numpy(),
                            x2_hat_crop=o["x2_hat"][:, :, cy:cy + 48, cx:cx + 48].numpy(),
                            x1_hat_rowsum=o["x1_hat"].double().sum(dim=(1, 3)).numpy(),
                            x2_hat_rowsum=o["x2_hat"].double().sum(dim=(1, 3)).numpy())
        print(tag, "delta1", float((o["x1_hat"] - x1).abs().mean()), "delta2", float((o["x2_hat"] - x2).abs().mean()))


if __name__ == "__main__":
    main()
