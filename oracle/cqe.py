"""CPU restatement of MASIC's cross quality enhancement network (TEST INFRASTRUCTURE, torch-CPU fp32).

`OracleIndependentEN` restates `Independent_EN` (coremasic/mywork/MASIC.py:1436-1501) with its helper
blocks `Enhancement_Block` (:149-164), `ResidualBlock` / `conv3x3` (compressai/layers/layers.py:105-190) and
`mask2weights_EN` (:1411-1434).  Modules are created in the reference's order with the reference's
initialisers, so `torch.manual_seed(s)` + construction reproduces its random init and `state_dict()` has
the same 86 entries.  Pinned against the unmodified reference by tests/test_oracle_cqe_pinned.py
(fixtures: tests/golden/make_golden_cqe.py).  Only tests/, smoke() and bench.py's cpu_baseline may import this.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .hsic import _Holder, _conv, _seq, warp, warp_masks


def _conv3x3(cin, cout):
    return nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1)              # layers/layers.py:105-107


def _residual_block(ch):                                                        # layers/layers.py:160-173
    h = _Holder()
    h.conv1 = _conv3x3(ch, ch)
    h.conv2 = _conv3x3(ch, ch)
    return h


def _run_residual_block(h, x):                                                  # layers/layers.py:175-190
    out = F.leaky_relu(h.conv1(x), 0.01)
    out = F.leaky_relu(h.conv2(out), 0.01)
    return out + x


def _enhancement_block(ch):                                                     # MASIC.py:149-154
    h = _Holder()
    h.RB1 = _residual_block(ch)
    h.RB2 = _residual_block(ch)
    h.RB3 = _residual_block(ch)
    return h


def _run_enhancement_block(h, x):                                               # MASIC.py:156-164
    out = _run_residual_block(h.RB1, x)
    out = _run_residual_block(h.RB2, out)
    out = _run_residual_block(h.RB3, out)
    return out + x


def _mask2weights_en(Kw=2):                                                     # MASIC.py:1411-1427
    h = _Holder()
    h.maskconv = _seq(_conv(1, Kw, 3, 1), nn.ReLU(inplace=True), _conv(Kw, 2 * Kw, 3, 1), nn.ReLU(inplace=True),
                      _conv(2 * Kw, 2 * Kw, 3, 1), nn.ReLU(inplace=True), _conv(2 * Kw, Kw, 3, 1))
    return h


class OracleIndependentEN(nn.Module):
    """Restatement of Independent_EN (MASIC.py:1436-1501)."""

    def __init__(self):
        super().__init__()
        self.EBl1 = _enhancement_block(32)           # MASIC.py:1440-1449, creation order kept
        self.EBl2 = _enhancement_block(64)
        self.EBl3 = _enhancement_block(96)
        self.EBr1 = _enhancement_block(32)
        self.EBr2 = _enhancement_block(64)
        self.EBr3 = _enhancement_block(96)
        self.conv0 = _conv3x3(3, 32)
        self.conv1 = _conv3x3(6, 32)
        self.conv2 = _conv3x3(96, 3)
        self.mask2weights_unit = _mask2weights_en()

    @torch.no_grad()
    def forward(self, x1_hat: torch.Tensor, x2_hat: torch.Tensor, Hm: torch.Tensor, keep=None):
        k = keep if keep is not None else {}
        h_inv = torch.inverse(Hm)                                               # :1457
        mask_r, mask_l = warp_masks(x1_hat, Hm)                                 # :1458
        w_r = F.softmax(self.mask2weights_unit.maskconv(mask_r), dim=-3)        # :1459 (:1429-1434)
        w_l = F.softmax(self.mask2weights_unit.maskconv(mask_l), dim=-3)        # :1460
        x1_warp = warp(x1_hat, Hm)                                              # :1461
        x2_warp = warp(x2_hat, h_inv)                                           # :1464
        x1_conv = self.conv0(x1_hat)                                            # :1467-1468
        x2_conv = self.conv0(x2_hat)
        out1 = torch.cat((x2_warp * w_l[:, 0:1], x1_hat * w_l[:, 1:2]), dim=-3)   # :1470
        out2 = torch.cat((x1_warp * w_r[:, 0:1], x2_hat * w_r[:, 1:2]), dim=-3)   # :1471
        out1 = self.conv1(out1)                                                 # :1473-1474
        out2 = self.conv1(out2)
        out1 = _run_enhancement_block(self.EBl1, out1)                          # :1476-1477
        out2 = _run_enhancement_block(self.EBr1, out2)
        k.update(w_r=w_r, w_l=w_l, eb1_l=out1, eb1_r=out2)
        out1_warp = warp(out1, Hm)                                              # :1479
        out2_warp = warp(out2, h_inv)                                           # :1480
        out1 = torch.cat((out1 * w_l[:, 1:2], out2_warp * w_l[:, 0:1]), dim=-3)   # :1481
        out2 = torch.cat((out2 * w_r[:, 1:2], out1_warp * w_r[:, 0:1]), dim=-3)   # :1482
        out1 = _run_enhancement_block(self.EBl2, out1)                          # :1483-1484
        out2 = _run_enhancement_block(self.EBr2, out2)
        out1 = torch.cat((out1, x1_conv), dim=-3)                               # :1486-1487
        out2 = torch.cat((out2, x2_conv), dim=-3)
        out1 = _run_enhancement_block(self.EBl3, out1)                          # :1488-1489
        out2 = _run_enhancement_block(self.EBr3, out2)
        out1 = self.conv2(out1)                                                 # :1491-1492
        out2 = self.conv2(out2)
        return {"x1_hat": out1 + x1_hat, "x2_hat": out2 + x2_hat}               # :1495-1501
