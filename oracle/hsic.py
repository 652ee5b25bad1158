"""CPU restatement of MASIC's HSIC codec forward pass (TEST INFRASTRUCTURE, torch-CPU fp32).

`OracleHSIC` owns a parameter tree whose `state_dict()` has exactly the reference's 248
entries (names, shapes, dtypes, order of creation — so `torch.manual_seed(s)` followed by
construction reproduces the reference's random init bit for bit), and `forward()` restates
coremasic/mywork/MASIC.py:744-851 (eval mode) on top of oracle.entropy and the kornia 0.5.0
restatement in oracle/shims.  Pinned against the unmodified reference in
tests/test_oracle_pinned.py.
"""
from __future__ import annotations

import sys
from pathlib import Path
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import entropy as E

_SHIMS = str(Path(__file__).resolve().parent / "shims")


def _kornia():
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    import kornia  # the restatement in oracle/shims (or the real package if one is installed first)
    return kornia


def warp(src: torch.Tensor, M: torch.Tensor) -> torch.Tensor:
    """kornia.warp_perspective(src, M, (H, W)) as called at MASIC.py:781,821,833."""
    return _kornia().warp_perspective(src, M, (src.shape[-2], src.shape[-1]))


def warp_masks(x: torch.Tensor, M: torch.Tensor):
    """MASIC.py:627-649 `mask()`: ones -> warp(M) -> warp(M^-1); the torch.where results are
    discarded there, so both masks stay fractional."""
    ones = torch.ones(x.shape[0], 1, x.shape[-2], x.shape[-1], dtype=x.dtype, device=x.device)
    m_r = warp(ones, M)
    m_l = warp(m_r, torch.inverse(M))
    return m_r, m_l


# ------------------------------------------------------------------ GDN (layers/gdn.py:41-92)
_PEDESTAL = (2.0 ** -18) ** 2


def nonneg_init(x: torch.Tensor) -> torch.Tensor:
    """ops/parametrizers.py:58-59."""
    ped = torch.tensor([_PEDESTAL], dtype=x.dtype, device=x.device)
    return torch.sqrt(torch.max(x + ped, ped))


def nonneg_forward(x: torch.Tensor, minimum: float) -> torch.Tensor:
    """ops/parametrizers.py:61-64: max(x, sqrt(minimum + pedestal))^2 - pedestal (fp32 buffers)."""
    bound = torch.tensor([(minimum + _PEDESTAL) ** 0.5], dtype=torch.float32, device=x.device)
    ped = torch.tensor([_PEDESTAL], dtype=torch.float32, device=x.device)
    return torch.max(x, bound) ** 2 - ped


def gdn(x: torch.Tensor, beta: torch.Tensor, gamma: torch.Tensor, inverse: bool,
        beta_min: float = 1e-6) -> torch.Tensor:
    """layers/gdn.py:77-92: x * rsqrt(beta' + gamma' (*) x^2)  (inverse: * sqrt)."""
    c = x.shape[1]
    b = nonneg_forward(beta, beta_min)
    g = nonneg_forward(gamma, 0.0).reshape(c, c, 1, 1)
    norm = F.conv2d(x ** 2, g, b)
    return x * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))


class _GDNParams(nn.Module):
    """Parameter/buffer holder with the reference GDN's state_dict layout (gdn.py:54-75)."""

    def __init__(self, c: int, inverse: bool = False, beta_min: float = 1e-6, gamma_init: float = 0.1):
        super().__init__()
        self.inverse = inverse
        self.beta_min = beta_min
        self.beta_reparam = _Reparam(beta_min)
        self.beta = nn.Parameter(nonneg_init(torch.ones(c)))
        self.gamma_reparam = _Reparam(0.0)
        self.gamma = nn.Parameter(nonneg_init(gamma_init * torch.eye(c)))

    def forward(self, x):
        return gdn(x, self.beta, self.gamma, self.inverse, self.beta_min)


class _Bound(nn.Module):
    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))


class _Reparam(nn.Module):
    def __init__(self, minimum: float):
        super().__init__()
        self.register_buffer("pedestal", torch.Tensor([_PEDESTAL]))
        self.lower_bound = _Bound((minimum + _PEDESTAL) ** 0.5)


# ------------------------------------------------------------------ entropy-model state holders
class _EBParams(nn.Module):
    """entropy_models.py:253-296 state (+ the EntropyModel buffers :78-83)."""

    def __init__(self, channels: int, init_scale: float = 10.0, filters=(3, 3, 3, 3), tail_mass: float = 1e-9):
        super().__init__()
        import numpy as np
        self.likelihood_lower_bound = _Bound(1e-9)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self._biases = nn.ParameterList()
        self._factors = nn.ParameterList()
        self._matrices = nn.ParameterList()
        f = (1,) + tuple(filters) + (1,)
        scale = init_scale ** (1 / (len(filters) + 1))
        for i in range(len(filters) + 1):
            init = np.log(np.expm1(1 / scale / f[i + 1]))
            m = torch.Tensor(channels, f[i + 1], f[i])
            m.data.fill_(init)
            self._matrices.append(nn.Parameter(m))
            b = torch.Tensor(channels, f[i + 1], 1)
            nn.init.uniform_(b, -0.5, 0.5)
            self._biases.append(nn.Parameter(b))
            if i < len(filters):
                fac = torch.Tensor(channels, f[i + 1], 1)
                nn.init.zeros_(fac)
                self._factors.append(nn.Parameter(fac))
        self.quantiles = nn.Parameter(torch.Tensor([-init_scale, 0, init_scale]).repeat(channels, 1, 1))
        t = np.log(2 / tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-t, 0, t]))

    def params(self) -> E.EBParams:
        return E.EBParams(list(self._matrices), list(self._biases), list(self._factors), self.quantiles)


class _GMMParams(nn.Module):
    """entropy_models.py:722-756 buffers of GaussianMixtureConditional_gf."""

    def __init__(self, K: int):
        super().__init__()
        self.K = K
        self.likelihood_lower_bound = _Bound(1e-9)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self.register_buffer("scale_table", torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([0.11]))
        self.lower_bound_scale = _Bound(0.11)


# ------------------------------------------------------------------ layer factories
def _conv(cin, cout, k=5, s=2):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=s, padding=k // 2)           # models/utils.py:128-135


def _deconv(cin, cout, k=5, s=2):
    return nn.ConvTranspose2d(cin, cout, kernel_size=k, stride=s, output_padding=s - 1, padding=k // 2)  # :138-146


class _MaskedConv(nn.Conv2d):
    """layers/layers.py:52-78, mask type 'A'; the reference multiplies weight.data by the mask
    in place on every forward — `apply_mask()` restates that side effect."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.register_buffer("mask", torch.ones_like(self.weight.data))
        _, _, h, w = self.mask.size()
        self.mask[:, :, h // 2, w // 2:] = 0
        self.mask[:, :, h // 2 + 1:] = 0

    def forward(self, x):
        self.weight.data *= self.mask
        return super().forward(x)


def _seq(*mods):
    return nn.Sequential(*mods)


class _Holder(nn.Module):
    pass


def _encoder(N, M, with_pre: bool):
    h = _Holder()
    if with_pre:
        h.pre_conv = _conv(6, 3, s=1)          # MASIC.py:559-560
        h.pre_gdn = _GDNParams(3)
    h.g_a_conv1 = _conv(3, N); h.g_a_gdn1 = _GDNParams(N)      # MASIC.py:513-519 / 562-568
    h.g_a_conv2 = _conv(N, N); h.g_a_gdn2 = _GDNParams(N)
    h.g_a_conv3 = _conv(N, N); h.g_a_gdn3 = _GDNParams(N)
    h.g_a_conv4 = _conv(N, M)
    return h


def _decoder(N, M, with_after: bool):
    h = _Holder()
    h.g_s_conv1 = _deconv(M, N); h.g_s_gdn1 = _GDNParams(N, inverse=True)   # MASIC.py:536-542 / 590-596
    h.g_s_conv2 = _deconv(N, N); h.g_s_gdn2 = _GDNParams(N, inverse=True)
    h.g_s_conv3 = _deconv(N, N); h.g_s_gdn3 = _GDNParams(N, inverse=True)
    h.g_s_conv4 = _deconv(N, 3)
    if with_after:
        h.after_gdn = _GDNParams(3, inverse=True)   # MASIC.py:599-600
        h.after_conv = _deconv(6, 3, s=1)
    return h


def _run_encoder(h, x):
    x = h.g_a_gdn1(h.g_a_conv1(x))
    x = h.g_a_gdn2(h.g_a_conv2(x))
    x = h.g_a_gdn3(h.g_a_conv3(x))
    return h.g_a_conv4(x)


def _run_decoder(h, y):
    y = h.g_s_gdn1(h.g_s_conv1(y))
    y = h.g_s_gdn2(h.g_s_conv2(y))
    y = h.g_s_gdn3(h.g_s_conv3(y))
    return h.g_s_conv4(y)


def _hyper_analysis(N, M):
    h = _Holder()
    h.encode_hyper = _seq(_conv(M, N, 5, 1), nn.ReLU(inplace=True), _conv(N, N, 5), nn.ReLU(inplace=True),
                          _conv(N, N, 5))                                  # MASIC.py:173-183
    return h


def _hyper_up(N, M):
    return _seq(_deconv(N, M, 5, 2), nn.LeakyReLU(inplace=True), _deconv(M, M * 3 // 2, 5, 2),
                nn.LeakyReLU(inplace=True), _conv(M * 3 // 2, M * 2, 3, 1))  # MASIC.py:678-691


def _gmm_net(N, M, K, first_is_deconv: bool, cin: int):
    """MASIC.py:330-376 (y1: 1x1 ConvTranspose2d for the first two layers) / :399-444 (y2)."""
    mk = _deconv if first_is_deconv else _conv
    h = _Holder()
    h.N, h.M, h.K = N, M, K
    h.gmm_sigma = _seq(mk(cin, 6 * M, 1, 1), nn.ReLU(inplace=True), mk(6 * M, 4 * M, 1, 1), nn.ReLU(inplace=True),
                       _conv(4 * M, M * K, 1, 1), nn.ReLU(inplace=True))
    h.gmm_means = _seq(mk(cin, 6 * M, 1, 1), nn.LeakyReLU(inplace=True), mk(6 * M, 4 * M, 1, 1),
                       nn.LeakyReLU(inplace=True), _conv(4 * M, M * K, 1, 1))
    h.gmm_weights = _seq(mk(cin, 6 * M, 1, 1), nn.LeakyReLU(inplace=True), mk(6 * M, M * K, 1, 1),
                         nn.LeakyReLU(inplace=True), _conv(M * K, M * K, 1, 1))
    return h


def _run_gmm_net(h, x):
    """MASIC.py:378-396 / :446-468: softmax over K on the (K, M) view of the weight logits."""
    sigma = h.gmm_sigma(x)
    mu = h.gmm_means(x)
    lw = h.gmm_weights(x)
    t = lw.reshape(-1, h.K, h.M, x.shape[-2], x.shape[-1])
    w = F.softmax(t, dim=-4).reshape(-1, h.M * h.K, x.shape[-2], x.shape[-1])
    return sigma, mu, w


def _mask2weights(Kw=3):
    h = _Holder()
    h.maskconv = _seq(_conv(1, 3, 3, 2), nn.ReLU(inplace=True), _conv(3, 6, 3), nn.ReLU(inplace=True),
                      _conv(6, 6, 3), nn.ReLU(inplace=True), _conv(6, 3, 3))   # MASIC.py:475-488
    h.Kw = Kw
    return h


def _run_mask2weights(h, m):
    o = h.maskconv(m)                                                       # MASIC.py:493-506
    t = o.reshape(-1, h.Kw, 1, o.shape[-2], o.shape[-1])
    return F.softmax(t, dim=-4).reshape(-1, h.Kw, o.shape[-2], o.shape[-1])


class OracleHSIC(nn.Module):
    """Restatement of HSIC (MASIC.py:652-851).  Module creation order == the reference's, so a
    seeded construction draws identical parameters."""

    def __init__(self, N: int = 128, M: int = 192, K: int = 5):
        super().__init__()
        self.N, self.M, self.K = N, M, K
        self.entropy_bottleneck1 = _EBParams(N)          # MASIC.py:50-54
        self.entropy_bottleneck2 = _EBParams(N)
        self.gaussian1 = _GMMParams(K)                   # :658-659
        self.gaussian2 = _GMMParams(K)
        self.encoder1 = _encoder(N, M, False)            # :664-667
        self.encoder2 = _encoder(N, M, True)
        self.decoder1 = _decoder(N, M, False)
        self.decoder2 = _decoder(N, M, True)
        self._h_a1 = _hyper_analysis(N, M)               # :672-673
        self._h_a2 = _hyper_analysis(N, M)
        self.h_s1_up = _hyper_up(N, M)                   # :678-691
        self.h_s2_up = _hyper_up(N, M)
        self.context_prediction1 = _MaskedConv(M, 2 * M, kernel_size=5, padding=2, stride=1)   # :693-703
        self.context_prediction2 = _MaskedConv(M, 2 * M, kernel_size=5, padding=2, stride=1)
        self._h_s1_same_resolution = _gmm_net(N, M, K, True, 4 * M)    # :704-705
        self._h_s2_same_resolution = _gmm_net(N, M, K, False, 5 * M)
        self.mask2weights_unit = _mask2weights(3)        # :706

    @torch.no_grad()
    def forward(self, x1: torch.Tensor, x2: torch.Tensor, Hm: torch.Tensor,
                keep: Optional[Dict[str, torch.Tensor]] = None):
        """Eval-mode MASIC.py:744-851.  `keep` (optional dict) receives intermediates."""
        K = self.K
        k = keep if keep is not None else {}
        # ---- left view
        y1 = _run_encoder(self.encoder1, x1)                                   # :746
        z1 = self._h_a1.encode_hyper(torch.abs(y1))                            # :747 (:184-187)
        z1_hat, z1_lik = E.eb_forward(self.entropy_bottleneck1.params(), z1)   # :749
        params1 = self.h_s1_up(z1_hat)                                         # :754
        y1_hat = torch.round(y1)                                               # :755
        ctx1 = self.context_prediction1(y1_hat)                                # :757
        s1, m1, w1 = _run_gmm_net(self._h_s1_same_resolution, torch.cat((params1, ctx1), dim=1))   # :765
        y1_hat, y1_lik = E.gmm_forward(y1, s1, m1, w1, K)                      # :767
        x1_hat = _run_decoder(self.decoder1, y1_hat)                           # :777
        # ---- right view
        x1_warp = warp(x1, Hm)                                                 # :781
        e2 = self.encoder2
        pre = e2.pre_gdn(e2.pre_conv(torch.cat((x1_warp, x2), dim=-3)))        # :573-574
        y2 = _run_encoder(e2, pre)                                             # :782
        z2 = self._h_a2.encode_hyper(torch.abs(y2))                            # :786
        z2_hat, z2_lik = E.eb_forward(self.entropy_bottleneck2.params(), z2)   # :787
        params2 = self.h_s2_up(z2_hat)                                         # :793
        y2_hat = torch.round(y2)                                               # :794
        ctx2 = self.context_prediction2(y2_hat)                                # :796
        mask_r, mask_l = warp_masks(x1, Hm)                                    # :803
        mw = _run_mask2weights(self.mask2weights_unit, mask_r)                 # :805
        x1_hat_warp = warp(x1_hat, Hm)                                         # :821 (== :833)
        y1w_hat = torch.round(_run_encoder(self.encoder1, x1_hat_warp))        # :822-824
        fused = torch.cat((params2 * mw[:, 0:1], ctx2 * mw[:, 1:2], y1w_hat * mw[:, 2:3]), dim=1)   # :827
        s2, m2, w2 = _run_gmm_net(self._h_s2_same_resolution, fused)
        y2_hat, y2_lik = E.gmm_forward(y2, s2, m2, w2, K)                      # :829
        d2 = self.decoder2
        core = _run_decoder(d2, y2_hat)                                        # :607-613
        x2_hat = d2.after_conv(torch.cat((d2.after_gdn(core), x1_hat_warp), dim=-3))   # :615-616
        k.update(y1=y1, z1=z1, y2=y2, z2=z2, params1=params1, ctx1=ctx1, sigma1=s1, mu1=m1, w1=w1,
                 params2=params2, ctx2=ctx2, sigma2=s2, mu2=m2, w2=w2, mask_weights=mw,
                 x1_warp=x1_warp, x1_hat_warp=x1_hat_warp, y1w_hat=y1w_hat, y2_hat=y2_hat, z2_hat=z2_hat)
        return {
            "x1_hat": x1_hat, "x2_hat": x2_hat, "y1_hat": y1_hat, "z1_hat": z1_hat,
            "x1_mask_R": mask_r, "x1_mask_L": mask_l,
            "likelihoods": {"y1": y1_lik, "y2": y2_lik, "z1": z1_lik, "z2": z2_lik},
        }


def bpp_of(out, num_pixels: int) -> float:
    """test2_real.py:88-100 / MASIC.py:126-128: sum over all four likelihood tensors."""
    import math
    return float(sum(torch.log(v).sum() / (-math.log(2) * num_pixels) for v in out["likelihoods"].values()))


def psnr_of(a: torch.Tensor, b: torch.Tensor) -> float:
    import math
    mse = torch.mean((a - b) ** 2).item()
    return 10.0 * math.log10(1.0 / mse) if mse > 0 else float("inf")


def synthetic_homography(batch: int, seed: int = 1) -> torch.Tensor:
    """identity + translation tx in [8,40], ty in [-6,6] + small shear/perspective."""
    g = torch.Generator().manual_seed(seed)
    Hm = torch.eye(3).repeat(batch, 1, 1)
    Hm[:, 0, 2] = 8 + 32 * torch.rand(batch, generator=g)
    Hm[:, 1, 2] = -6 + 12 * torch.rand(batch, generator=g)
    Hm[:, 0, 1] = 1e-2 * (2 * torch.rand(batch, generator=g) - 1)
    Hm[:, 2, 0] = 1e-6 * (2 * torch.rand(batch, generator=g) - 1)
    return Hm
