"""Bring a random-init HSIC into the regime the codec is USED in (reconstructions tens of dB above random init) with
the repo's own CUDA training step, on smooth synthetic stereo pairs, so that the parity tolerances (|dPSNR| <= 0.01 dB,
|dbpp| <= 0.1 %) can be checked where they are not vacuous.

    python tools/train_regime.py [--steps 300] [--lr 3e-4] [--clip 1.0] [--lmbda 0.05] [--size 256 256] [--eval-size 256 384]

`train_to_psnr` is what tests/test_hsic_trained_gpu.py calls; run as a script it prints the trajectory and the
CUDA-vs-oracle comparison of the trained weights."""
import argparse
import math
import sys
import time
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def smooth_pairs(batch, h, w, gen, shift_px=None):
    """Low-frequency colour fields in [0,1] (sum of bicubically upsampled noise at three scales); the right view is
    the left one shifted by `tx` pixels (plus a little independent detail), H = that translation."""
    def field(hh, ww):
        f = 0.5 + sum(a * F.interpolate(torch.randn(batch, 3, max(2, hh // s), max(2, ww // s), generator=gen),
                                        size=(hh, ww), mode="bicubic", align_corners=False)
                      for s, a in ((64, 0.22), (16, 0.08), (8, 0.03)))
        return f
    tx = 16 if shift_px is None else shift_px
    wide = field(h, w + tx).clamp(0.0, 1.0)
    x1 = wide[..., tx:].contiguous()
    x2 = (wide[..., :w] + 0.01 * F.interpolate(torch.randn(batch, 3, h // 8, w // 8, generator=gen), size=(h, w),
                                               mode="bicubic", align_corners=False)).clamp(0.0, 1.0).contiguous()
    Hm = torch.eye(3).repeat(batch, 1, 1)
    Hm[:, 0, 2] = float(tx)          # x1 warped by H lands on x2: dst(u) = src(u - tx)
    return x1, x2, Hm


def train_to_psnr(net, dev, *, target_db=25.0, max_steps=600, batch=2, size=(256, 256), lr=3e-4, lmbda=0.05,
                  seed=0, check_every=50, clip=1.0, log=None):
    """Adam on net.parameters() (+ the aux optimiser) with HSICTrainer.train_step until the training PSNR (from the
    step's own mse) reaches target_db.  Returns (steps, last psnr)."""
    h, w = size
    net.train()
    tr = net.trainer(batch, h, w, dev, lmbda=lmbda)
    opt = torch.optim.Adam(net.parameters(), lr=lr)
    aux = torch.optim.Adam(net.aux_parameters(), lr=1e-3)
    gen = torch.Generator().manual_seed(seed)
    psnr, hist = 0.0, []
    for step in range(1, max_steps + 1):
        x1, x2, Hm = smooth_pairs(batch, h, w, gen)
        res = tr.train_step(x1.to(dev), x2.to(dev), Hm.to(dev), opt, aux, clip_max_norm=clip)
        if not math.isfinite(res["loss"]):
            raise RuntimeError(f"training diverged at step {step}: {res}")
        hist.append(res["mse"] / 2)
        if step % check_every == 0:
            m = sum(hist[-check_every:]) / check_every
            psnr = 10 * math.log10(1.0 / m)
            if log:
                log(f"step {step}: loss {res['loss']:.3f} bpp {res['bpp']:.3f} psnr(train, mean of two views) {psnr:.2f} dB")
            if psnr >= target_db:
                break
    net.eval()
    net.invalidate_engines()
    del tr
    return step, psnr


def compare_with_oracle(net, dev, h, w, seed=9):
    from oracle import hsic as OH
    oracle = OH.OracleHSIC(128, 192, 5).eval()
    oracle.load_state_dict({k: v.detach().cpu() for k, v in net.state_dict().items()})
    gen = torch.Generator().manual_seed(seed)
    x1, x2, Hm = smooth_pairs(1, h, w, gen)
    ref = oracle(x1, x2, Hm)
    with torch.no_grad():
        out = net(x1.to(dev), x2.to(dev), Hm.to(dev))
    cpu = lambda t: t.detach().cpu()   # noqa: E731
    npx = h * w
    bpp = lambda o: sum(float(torch.log(cpu(v).double()).sum() / (-math.log(2) * npx)) for v in o["likelihoods"].values())  # noqa: E731
    psnr = lambda a, b: 10 * math.log10(1.0 / float(torch.mean((cpu(a).double() - b.double()) ** 2)))  # noqa: E731
    r = dict(shape=[1, h, w], bpp_oracle=bpp(ref), bpp_cuda=bpp(out),
             psnr1_oracle=psnr(ref["x1_hat"], x1), psnr1_cuda=psnr(out["x1_hat"], x1),
             psnr2_oracle=psnr(ref["x2_hat"], x2), psnr2_cuda=psnr(out["x2_hat"], x2),
             y1_symbol_flips=float((cpu(out["y1_hat"]) != ref["y1_hat"]).float().mean()),
             y1_nonzero=float((ref["y1_hat"] != 0).float().mean()),
             xhat1_rms_diff=float((cpu(out["x1_hat"]) - ref["x1_hat"]).pow(2).mean().sqrt()),
             xhat2_rms_diff=float((cpu(out["x2_hat"]) - ref["x2_hat"]).pow(2).mean().sqrt()))
    r["dbpp_rel"] = abs(r["bpp_cuda"] - r["bpp_oracle"]) / r["bpp_oracle"]
    r["dpsnr1_db"] = abs(r["psnr1_cuda"] - r["psnr1_oracle"])
    r["dpsnr2_db"] = abs(r["psnr2_cuda"] - r["psnr2_oracle"])
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--clip", type=float, default=1.0)
    ap.add_argument("--lmbda", type=float, default=0.05)
    ap.add_argument("--target", type=float, default=99.0)
    ap.add_argument("--size", type=int, nargs=2, default=[256, 256])
    ap.add_argument("--eval-size", type=int, nargs=2, default=[256, 384])
    ap.add_argument("--save-left", default="", help="save the left-view chain's weights (bf16) for CPU-side analysis")
    a = ap.parse_args()
    from masic_b200.hsic import HSIC
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = HSIC().to(dev)
    t0 = time.perf_counter()
    steps, psnr = train_to_psnr(net, dev, target_db=a.target, max_steps=a.steps, size=tuple(a.size), lr=a.lr,
                                lmbda=a.lmbda, clip=a.clip, log=print)
    print(f"trained {steps} steps in {time.perf_counter() - t0:.1f} s, train psnr {psnr:.2f} dB (lr {a.lr}, lambda {a.lmbda})")
    print(compare_with_oracle(net, dev, *a.eval_size))
    if a.save_left:
        keep = ("encoder1.", "decoder1.", "_h_a1.", "h_s1_up.", "context_prediction1.", "_h_s1_same_resolution.",
                "entropy_bottleneck1.", "gaussian1.")
        sd = {k: (v.detach().cpu().to(torch.bfloat16) if v.is_floating_point() and v.numel() > 4096 else v.detach().cpu())
              for k, v in net.state_dict().items() if k.startswith(keep)}
        torch.save(sd, a.save_left)
        print("saved", a.save_left, sum(v.numel() for v in sd.values()), "values")


if __name__ == "__main__":
    main()
