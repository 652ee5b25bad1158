"""Where does the CUDA engine's deviation from the fp32 oracle come from?  CPU-side simulation of the engine's roundings
(bf16 weights, bf16 inter-layer activations, bf16 x^2 / gamma' in the fused GDN) on the LEFT-view chain of a trained
model (weights saved by `tools/train_regime.py --save-left`), switching individual roundings on and off.

    python tools/noise_analysis.py gpurun_out/trained_left.pt [H W]

Prints, per variant, the rms deviation of x1_hat from the exact fp32 result, the PSNR change it causes, and the share
of latent symbols that flip.  (Test infrastructure: uses oracle/.)"""
import math
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import hsic as OH  # noqa: E402
from tools.train_regime import smooth_pairs  # noqa: E402


def bf(t):
    return t.to(torch.bfloat16).float()


def hf(t):
    return t.to(torch.float16).float()


class Sim:
    """flags: w (weights bf16), a (layer inputs bf16), g (GDN x^2 and gamma' bf16), per-layer overrides in `exact`."""

    def __init__(self, w=True, a=True, g=True, exact=(), fmt=bf, g_x2=True, g_gamma=True, gamma_fmt=bf):
        self.w, self.a, self.g, self.exact = w, a, g, set(exact)
        self.fmt, self.g_x2, self.g_gamma, self.gamma_fmt = fmt, g_x2, g_gamma, gamma_fmt

    def conv(self, name, mod, x, transposed):
        ex = name in self.exact
        wt = mod.weight if (ex or not self.w) else self.fmt(mod.weight)
        xi = x if (ex or not self.a) else self.fmt(x)
        if transposed:
            return F.conv_transpose2d(xi, wt, mod.bias, stride=mod.stride, padding=mod.padding, output_padding=mod.output_padding)
        return F.conv2d(xi, wt, mod.bias, stride=mod.stride, padding=mod.padding)

    def gdn(self, name, mod, x):
        c = x.shape[1]
        b = OH.nonneg_forward(mod.beta, mod.beta_min)
        g = OH.nonneg_forward(mod.gamma, 0.0).reshape(c, c, 1, 1)
        x2 = x ** 2
        if self.g and name not in self.exact:
            if self.g_x2:
                x2 = bf(x2)
            if self.g_gamma:
                g = self.gamma_fmt(g)
        norm = F.conv2d(x2, g, b)
        return x * (torch.sqrt(norm) if mod.inverse else torch.rsqrt(norm))


def run(net, x1, sim, y_hat_override=None):
    e, d = net.encoder1, net.decoder1
    with torch.no_grad():
        x = x1
        for i in (1, 2, 3):
            x = sim.gdn(f"gdn{i}", getattr(e, f"g_a_gdn{i}"), sim.conv(f"conv{i}", getattr(e, f"g_a_conv{i}"), x, False))
        y = sim.conv("conv4", e.g_a_conv4, x, False)
        y_hat = torch.round(y) if y_hat_override is None else y_hat_override
        t = y_hat
        for i in (1, 2, 3):
            t = sim.gdn(f"igdn{i}", getattr(d, f"g_s_gdn{i}"), sim.conv(f"deconv{i}", getattr(d, f"g_s_conv{i}"), t, True))
        x_hat = sim.conv("deconv4", d.g_s_conv4, t, True)
    return y, y_hat, x_hat


def main():
    path = sys.argv[1]
    h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (256, 384)
    sd = torch.load(path)
    net = OH.OracleHSIC(128, 192, 5).eval()
    full = net.state_dict()
    for k, v in sd.items():
        full[k] = v.float() if v.is_floating_point() else v
    net.load_state_dict(full)
    g = torch.Generator().manual_seed(9)
    x1, _, _ = smooth_pairs(1, h, w, g)
    exact = Sim(w=False, a=False, g=False)
    y0, yh0, xh0 = run(net, x1, exact)
    mse0 = float(((xh0 - x1) ** 2).mean())
    print(f"exact fp32: PSNR {10 * math.log10(1 / mse0):.3f} dB, |y|max {float(y0.abs().max()):.1f}, nonzero {float((yh0 != 0).float().mean()):.3f}")

    def report(name, sim, same_symbols=False):
        y, yh, xh = run(net, x1, sim, y_hat_override=yh0 if same_symbols else None)
        mse = float(((xh - x1) ** 2).mean())
        rms = float(((xh - xh0) ** 2).mean().sqrt())
        flips = float((yh != yh0).float().mean())
        print(f"{name:58s} rms(x_hat dev) {rms:.3e}  dPSNR {10 * math.log10(mse0 / mse):+.5f} dB  flips {flips:.5f}  "
              f"rms(y dev) {float(((y - y0) ** 2).mean().sqrt()):.3e}")

    print("---- the inference engine of round 2: fp16 weights / activations, bf16 x^2 and gamma' in the fused GDN")
    report("fp16 w, a + bf16 gdn operands", Sim(fmt=hf))
    report("  ... decoder only (exact symbols fed)", Sim(fmt=hf), same_symbols=True)
    report("fp16 w, a, exact gdn", Sim(fmt=hf, g=False))
    report("  ... decoder only", Sim(fmt=hf, g=False), same_symbols=True)
    report("only gdn operands bf16 (x^2 and gamma')", Sim(w=False, a=False), same_symbols=True)
    report("only gamma' bf16", Sim(w=False, a=False, g_x2=False), same_symbols=True)
    report("only x^2 bf16", Sim(w=False, a=False, g_gamma=False), same_symbols=True)
    report("only gamma' fp16", Sim(w=False, a=False, g_x2=False, gamma_fmt=hf), same_symbols=True)
    report("fp16 w, a + bf16 x^2 + fp16 gamma'", Sim(fmt=hf, gamma_fmt=hf))
    report("  ... decoder only", Sim(fmt=hf, gamma_fmt=hf), same_symbols=True)
    print("---- round 1 numerics (everything bf16)")
    report("engine numerics (w, a, gdn in bf16)", Sim())
    report("  ... decoder only (exact symbols fed)", Sim(), same_symbols=True)
    report("weights bf16 only", Sim(a=False, g=False))
    report("weights bf16 only, decoder only", Sim(a=False, g=False), same_symbols=True)
    report("activations bf16 only", Sim(w=False, g=False))
    report("activations bf16 only, decoder only", Sim(w=False, g=False), same_symbols=True)
    report("gdn x^2/gamma bf16 only", Sim(w=False, a=False))
    report("gdn x^2/gamma bf16 only, decoder only", Sim(w=False, a=False), same_symbols=True)
    for lay in ("deconv4", "deconv3", "igdn3", "deconv2", "igdn2", "deconv1", "igdn1"):
        report(f"all bf16 except {lay} exact, decoder only", Sim(exact=(lay,)), same_symbols=True)
    report("all bf16 except deconv3+deconv4+igdn3 exact, decoder only", Sim(exact=("deconv3", "deconv4", "igdn3")), same_symbols=True)
    for lay in ("conv1", "conv2", "conv3", "conv4", "gdn1", "gdn2", "gdn3"):
        report(f"all bf16 except {lay} exact (encoder side)", Sim(exact=(lay,)))


if __name__ == "__main__":
    main()
