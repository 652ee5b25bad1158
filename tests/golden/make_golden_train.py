"""Golden fixture of ONE TRAINING STEP of the unmodified reference (tests/golden/hsic_train_*.npz).

    make -C oracle ref && python tests/golden/make_golden_train.py

`MASIC.HSIC(...).train()` from /root/reference runs forward -> RateDistortionLoss -> backward -> aux loss ->
backward exactly as newtrain_codec_real.py:105-146 does.  The only intervention: the reference draws its
quantisation noise inside `EntropyModel._get_noise_cached` (entropy_models.py:88-96); for a reproducible step
that method is replaced AT RUN TIME (no reference file is edited) by one that hands out the seeded tensors of
oracle.train.make_noise in call order.  Stored: loss / bpp / mse / aux loss, and for every parameter the
gradient's L2 norm, sum, and its first 48 values — enough to pin oracle/train.py, which the GPU parity tests
then differentiate on the box.
"""
from __future__ import annotations

import math
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
from oracle import refimport  # noqa: E402
from oracle import train as OT  # noqa: E402
from oracle.hsic import OracleHSIC, synthetic_homography  # noqa: E402

torch.set_num_threads(8)
LMBDA = 0.001


def run_case(MASIC, tag, batch, h, w, scale_y, noise_seed=77):
    from compressai.entropy_models import entropy_models as EM
    torch.manual_seed(0)
    net = MASIC.HSIC(N=128, M=192, K=5).train()
    if scale_y != 1.0:
        with torch.no_grad():
            net.encoder1.g_a_conv4.weight.mul_(scale_y)
            net.encoder2.g_a_conv4.weight.mul_(scale_y)
    g = torch.Generator().manual_seed(100)
    x1 = torch.rand(batch, 3, h, w, generator=g)
    x2 = torch.rand(batch, 3, h, w, generator=g)
    Hm = synthetic_homography(batch, seed=1)
    shapes = OracleHSIC.__new__(OracleHSIC)
    shapes.M, shapes.N = 192, 128
    noise = OT.make_noise(shapes, batch, h, w, noise_seed)
    queue = [noise[k] for k in OT.NOISE_KEYS]

    def fake_noise(self, x):
        t = queue.pop(0)
        if tuple(x.shape) != tuple(t.shape):            # EntropyBottleneck works on the (C,1,N*H*W) view
            n, c, hh, ww = t.shape
            t = t.permute(1, 2, 3, 0).reshape(c, 1, -1)
        assert tuple(t.shape) == tuple(x.shape), (t.shape, x.shape)
        return t

    orig = EM.EntropyModel._get_noise_cached
    EM.EntropyModel._get_noise_cached = fake_noise
    try:
        out = net(x1, x2, Hm)
    finally:
        EM.EntropyModel._get_noise_cached = orig
    assert not queue
    num_pixels = batch * h * w
    bpp = sum(torch.log(l).sum() / (-math.log(2) * num_pixels) for l in out["likelihoods"].values())
    mse = torch.nn.functional.mse_loss(out["x1_hat"], x1) + torch.nn.functional.mse_loss(out["x2_hat"], x2)
    loss = LMBDA * 255 ** 2 * mse + bpp                         # newtrain_codec_real.py:79
    loss.backward()
    aux = net.aux_loss()
    aux.backward()
    names, norms, sums, heads = [], [], [], []
    for n, p in torch.nn.Module.named_parameters(net):
        gr = p.grad if p.grad is not None else torch.zeros_like(p)
        names.append(n)
        norms.append(float(gr.double().norm()))
        sums.append(float(gr.double().sum()))
        hd = torch.zeros(48)
        k = min(48, gr.numel())
        hd[:k] = gr.reshape(-1)[:k]
        heads.append(hd.numpy())
    np.savez_compressed(HERE / f"hsic_train_{tag}.npz", batch=batch, h=h, w=w, scale_y=scale_y, lmbda=LMBDA,
                        x_seed=100, h_seed=1, noise_seed=noise_seed, loss=float(loss), bpp=float(bpp), mse=float(mse),
                        aux=float(aux), names=np.array(names), grad_norm=np.array(norms), grad_sum=np.array(sums),
                        grad_head=np.stack(heads), x1_hat_mean=float(out["x1_hat"].mean()),
                        y1_hat_head=out["y1_hat"].detach().reshape(-1)[:64].numpy())
    print(tag, "loss", float(loss), "bpp", float(bpp), "mse", float(mse), "aux", float(aux),
          "nonzero grads", sum(1 for v in norms if v > 0), "/", len(norms))


def main():
    MASIC = refimport.import_masic()
    run_case(MASIC, "scaled8_b2_64x128", 2, 64, 128, 8.0)
    run_case(MASIC, "init_b1_128x192", 1, 128, 192, 1.0)


if __name__ == "__main__":
    main()
