"""TEST INFRASTRUCTURE — CPU restatement of MASIC's training step (forward in train() mode + loss),
differentiated by torch autograd.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import this; nothing under masic_b200/ does.

Follows (reference, file:line):
  HSIC.forward, training branch            coremasic/mywork/MASIC.py:744-851 ('noise' at :755,:794,:823)
  EntropyModel._quantize('noise')          compressai/entropy_models/entropy_models.py:98-110  (x + U(-.5,.5))
  EntropyBottleneck.forward / .loss        entropy_models.py:384-411, :345-348
  GaussianMixtureConditional_gf.forward    entropy_models.py:808-858  (second, independent noise draw)
  LowerBound custom gradient               compressai/ops/bound_ops.py:36-58
  NonNegativeParametrizer.forward          compressai/ops/parametrizers.py:61-64 (LowerBound inside)
  RateDistortionLoss                       coremasic/mywork/newtrain_codec_real.py:66-87

The reference draws fresh uniform noise inside the model (`_get_noise_cached`, entropy_models.py:88-96);
here the seven noise tensors are INPUTS (`noise` dict) so both sides of a parity test see the same draw.
Call order in the reference's forward = key order of NOISE_KEYS.

Pinned by tests/golden/hsic_train_*.npz (generated from the unmodified reference in train() mode with
`_get_noise_cached` returning the same tensors): loss, aux loss and every parameter gradient.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from . import entropy as E
from . import hsic as OH

NOISE_KEYS = ("z1", "y1_ctx", "y1", "z2", "y2_ctx", "y1w", "y2")


class LowerBoundFn(torch.autograd.Function):
    """bound_ops.py:36-58: max(x, bound); gradient passes where x >= bound or it pushes x up (grad < 0)."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, g):
        x, bound = ctx.saved_tensors
        return ((x >= bound) | (g < 0)) * g, None


def lower_bound(x: torch.Tensor, bound: float) -> torch.Tensor:
    return LowerBoundFn.apply(x, torch.tensor([bound], dtype=x.dtype))


def nonneg(x: torch.Tensor, minimum: float) -> torch.Tensor:
    """parametrizers.py:61-64 with the LowerBound gradient."""
    return lower_bound(x, (minimum + OH._PEDESTAL) ** 0.5) ** 2 - OH._PEDESTAL


def gdn(x, p: "OH._GDNParams"):
    c = x.shape[1]
    b = nonneg(p.beta, p.beta_min)
    g = nonneg(p.gamma, 0.0).reshape(c, c, 1, 1)
    norm = F.conv2d(x ** 2, g, b)
    return x * (torch.sqrt(norm) if p.inverse else torch.rsqrt(norm))


def run_encoder(h, x):
    x = gdn(h.g_a_conv1(x), h.g_a_gdn1)
    x = gdn(h.g_a_conv2(x), h.g_a_gdn2)
    x = gdn(h.g_a_conv3(x), h.g_a_gdn3)
    return h.g_a_conv4(x)


def run_decoder(h, y):
    y = gdn(h.g_s_conv1(y), h.g_s_gdn1)
    y = gdn(h.g_s_conv2(y), h.g_s_gdn2)
    y = gdn(h.g_s_conv3(y), h.g_s_gdn3)
    return h.g_s_conv4(y)


class _EBLive:
    """E.EBParams without the detach: the live parameters of an oracle bottleneck (autograd flows)."""

    def __init__(self, eb, detach_net: bool = False):
        d = (lambda t: t.detach()) if detach_net else (lambda t: t)
        self.matrices = [d(m) for m in eb._matrices]
        self.biases = [d(b) for b in eb._biases]
        self.factors = [d(f) for f in eb._factors]
        self.quantiles = eb.quantiles


def eb_forward_train(p, z: torch.Tensor, noise: torch.Tensor):
    """entropy_models.py:384-411, training: outputs = z + noise; likelihood floored with LowerBound(1e-9)."""
    zc = z.permute(1, 2, 3, 0).contiguous()
    shape = zc.shape
    vals = zc.reshape(shape[0], 1, -1) + noise.permute(1, 2, 3, 0).reshape(shape[0], 1, -1)
    lik = lower_bound(E.eb_likelihood(p, vals), E.LIKELIHOOD_BOUND)
    back = lambda t: t.reshape(shape).permute(3, 0, 1, 2).contiguous()  # noqa: E731
    return back(vals), back(lik)


def gmm_forward_train(y, sigma, mu, w, K: int, noise: torch.Tensor):
    """entropy_models.py:808-858, training: outputs = y + noise (means=None); scale floor 0.11 and
    likelihood floor 1e-9 both through LowerBound."""
    out = y + noise
    M = y.shape[1]
    lik = None
    for k in range(K):
        sl = slice(M * k, M * (k + 1))
        v = torch.abs(out - mu[:, sl])
        s = lower_bound(sigma[:, sl], E.SCALE_BOUND)
        term = (E.std_cumulative((0.5 - v) / s) - E.std_cumulative((-0.5 - v) / s)) * w[:, sl]
        lik = term if lik is None else lik + term
    return out, lower_bound(lik, E.LIKELIHOOD_BOUND)


def forward_train(net: "OH.OracleHSIC", x1, x2, Hm, noise: Dict[str, torch.Tensor]):
    """MASIC.py:744-851 with self.training == True (autograd enabled)."""
    K = net.K
    y1 = run_encoder(net.encoder1, x1)                                             # :746
    z1 = net._h_a1.encode_hyper(torch.abs(y1))                                     # :747
    z1_hat, z1_lik = eb_forward_train(_EBLive(net.entropy_bottleneck1), z1, noise["z1"])   # :749
    params1 = net.h_s1_up(z1_hat)                                                  # :754
    ctx1 = net.context_prediction1(y1 + noise["y1_ctx"])                           # :755-757
    s1, m1, w1 = OH._run_gmm_net(net._h_s1_same_resolution, torch.cat((params1, ctx1), dim=1))   # :765
    y1_hat, y1_lik = gmm_forward_train(y1, s1, m1, w1, K, noise["y1"])             # :767
    x1_hat = run_decoder(net.decoder1, y1_hat)                                     # :777
    x1_warp = OH.warp(x1, Hm)                                                      # :781
    e2 = net.encoder2
    pre = gdn(e2.pre_conv(torch.cat((x1_warp, x2), dim=-3)), e2.pre_gdn)           # :573-574
    y2 = run_encoder(e2, pre)                                                      # :782
    z2 = net._h_a2.encode_hyper(torch.abs(y2))                                     # :786
    z2_hat, z2_lik = eb_forward_train(_EBLive(net.entropy_bottleneck2), z2, noise["z2"])   # :787
    params2 = net.h_s2_up(z2_hat)                                                  # :793
    ctx2 = net.context_prediction2(y2 + noise["y2_ctx"])                           # :794-796
    mask_r, mask_l = OH.warp_masks(x1, Hm)                                         # :803
    mw = OH._run_mask2weights(net.mask2weights_unit, mask_r)                       # :805
    x1_hat_warp = OH.warp(x1_hat, Hm)                                              # :821 (== :833)
    y1w_hat = run_encoder(net.encoder1, x1_hat_warp) + noise["y1w"]                # :822-824
    fused = torch.cat((params2 * mw[:, 0:1], ctx2 * mw[:, 1:2], y1w_hat * mw[:, 2:3]), dim=1)   # :827
    s2, m2, w2 = OH._run_gmm_net(net._h_s2_same_resolution, fused)
    y2_hat, y2_lik = gmm_forward_train(y2, s2, m2, w2, K, noise["y2"])             # :829
    d2 = net.decoder2
    core = run_decoder(d2, y2_hat)                                                 # :607-613
    x2_hat = d2.after_conv(torch.cat((gdn(core, d2.after_gdn), x1_hat_warp), dim=-3))   # :615-616
    return {"x1_hat": x1_hat, "x2_hat": x2_hat, "y1_hat": y1_hat, "z1_hat": z1_hat,
            "x1_mask_R": mask_r, "x1_mask_L": mask_l,
            "likelihoods": {"y1": y1_lik, "y2": y2_lik, "z1": z1_lik, "z2": z2_lik}}


def rd_loss(out, x1, x2, lmbda: float):
    """newtrain_codec_real.py:66-87."""
    n, _, h, w = x1.shape
    num_pixels = n * h * w
    bpp = sum(torch.log(l).sum() / (-math.log(2) * num_pixels) for l in out["likelihoods"].values())
    mse = F.mse_loss(out["x1_hat"], x1) + F.mse_loss(out["x2_hat"], x2)
    return lmbda * 255 ** 2 * mse + bpp, bpp, mse


def aux_loss(net: "OH.OracleHSIC"):
    """MASIC.py:59-65 + entropy_models.py:345-348: sum over both bottlenecks of |logits(quantiles) - target|,
    matrices/biases/factors detached."""
    total = 0.0
    for eb in (net.entropy_bottleneck1, net.entropy_bottleneck2):
        det = _EBLive(eb, detach_net=True)
        total = total + torch.abs(E.eb_logits_cumulative(det, eb.quantiles) - eb.target).sum()
    return total


def make_noise(net: "OH.OracleHSIC", batch: int, h: int, w: int, seed: int) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    ys, zs = (batch, net.M, h // 16, w // 16), (batch, net.N, h // 64, w // 64)
    return {k: torch.rand(*(zs if k.startswith("z") else ys), generator=g) - 0.5 for k in NOISE_KEYS}


def train_step_grads(net: "OH.OracleHSIC", x1, x2, Hm, noise, lmbda: float):
    """One forward + backward of the main loss and of the aux loss; returns (loss, bpp, mse, aux, grads)
    where grads maps parameter name -> gradient (main loss; `quantiles` carry the aux-loss gradient, as the
    reference's two optimizers see them, newtrain_codec_real.py:134-146)."""
    net.train()
    for p in net.parameters():
        p.grad = None
    out = forward_train(net, x1, x2, Hm, noise)
    loss, bpp, mse = rd_loss(out, x1, x2, lmbda)
    loss.backward()
    aux = aux_loss(net)
    aux.backward()
    grads = {n: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for n, p in net.named_parameters()}
    return float(loss), float(bpp), float(mse), float(aux), grads, out
