"""Tensor-core weight-gradient kernel (masic_b200/csrc/wgrad_tc.cu) against torch autograd (fp32, same bf16-rounded
operands) on the layer geometries of HSIC: conv / transposed conv, k in {1,3,5}, stride {1,2}, masked taps,
channel slices of wider buffers, ragged spatial sizes."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    from masic_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")


def _nhwc(t, pitch, coff):
    n, c, h, w = t.shape
    buf = torch.randn(n, h, w, pitch, device=t.device).to(torch.bfloat16)      # garbage around the slice
    buf[..., coff:coff + c] = t.permute(0, 2, 3, 1).to(torch.bfloat16)
    return buf.contiguous()


CASES = [
    # name, transposed, k, stride, c_lo, c_hi, h_lo, w_lo, n, (lo_pitch, lo_coff), (hi_pitch, hi_coff), masked
    ("conv5s2_128", False, 5, 2, 128, 128, 24, 40, 1, (128, 0), (128, 0), False),
    ("conv5s2_192out_b2", False, 5, 2, 192, 128, 10, 14, 2, (192, 0), (128, 0), False),
    ("conv5s1_h_a", False, 5, 1, 128, 192, 20, 28, 1, (128, 0), (192, 0), False),
    ("conv3s1_pitch", False, 3, 1, 384, 320, 18, 24, 1, (384, 0), (384, 0), False),
    ("conv1x1_slices", False, 1, 1, 768, 1152, 12, 20, 1, (1536, 768), (3456, 1152), False),
    ("conv1x1_oddtiles", False, 1, 1, 960, 960, 9, 7, 2, (960, 0), (960, 0), False),
    ("masked5x5", False, 5, 1, 384, 192, 16, 24, 1, (768, 384), (192, 0), True),
    ("deconv5s2_128", True, 5, 2, 128, 128, 12, 20, 1, (128, 0), (128, 0), False),
    ("deconv5s2_192in", True, 5, 2, 192, 128, 8, 14, 2, (192, 0), (128, 0), False),
    # the 3-channel side of g_a_conv1 / g_s_conv4: the image travels as a dense 16-channel-pitch NHWC buffer
    ("conv5s2_img16", False, 5, 2, 128, 16, 24, 40, 2, (128, 0), (16, 0), False),
    ("conv5s2_img16_big", False, 5, 2, 128, 16, 64, 112, 2, (128, 0), (16, 0), False),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_wgrad_matches_autograd(dev, case):
    from masic_b200.convplan import MASK_A_5x5, WgradPlan
    name, transposed, k, s, c_lo, c_hi, h_lo, w_lo, n, (lp, lc), (hp, hc), masked = case
    g = torch.Generator(device="cpu").manual_seed(hash(name) % 1000)
    lo = torch.randn(n, c_lo, h_lo, w_lo, generator=g).to(dev)
    hi = torch.randn(n, c_hi, h_lo * s, w_lo * s, generator=g).to(dev)
    lo_b, hi_b = _nhwc(lo, lp, lc), _nhwc(hi, hp, hc)
    lo_r = lo_b[..., lc:lc + c_lo].float().permute(0, 3, 1, 2).contiguous()
    hi_r = hi_b[..., hc:hc + c_hi].float().permute(0, 3, 1, 2).contiguous()
    w = torch.zeros(c_lo, c_hi, k, k, device=dev, requires_grad=True)
    mask = torch.ones(k, k, device=dev)
    if masked:
        mask[2, 2:] = 0
        mask[3:] = 0
    # dW[cl][ch][k] = sum LO[p, cl] * HI[s*p + k - pad, ch]: the weight gradient of conv2d(HI, W) w.r.t. W with
    # output gradient LO (for a transposed conv the same sum with the roles of input and gradient exchanged)
    out = F.conv2d(hi_r, w * mask, stride=s, padding=k // 2)
    assert out.shape == lo_r.shape
    (out * lo_r).sum().backward()
    want = w.grad
    dw = torch.full((c_lo, c_hi, k, k), 7.0, device=dev)
    plan = WgradPlan(ksize=k, stride=s, lo=lo_b, c_lo=c_lo, lo_coff=lc, hi=hi_b, c_hi=c_hi, hi_coff=hc, dw=dw,
                     tap_mask=MASK_A_5x5 if masked else 0)
    plan.launch()
    torch.cuda.synchronize()
    if masked:
        assert bool((dw[:, :, mask == 0] == 7.0).all())          # dead taps untouched
        dw = dw * mask
    scale = float(want.abs().max())
    err = float((dw - want).abs().max())
    assert err <= 2e-3 * scale, (name, err, scale)
    # accumulate=1 adds a second pass on top (encoder1 runs twice per step)
    plan2 = WgradPlan(ksize=k, stride=s, lo=lo_b, c_lo=c_lo, lo_coff=lc, hi=hi_b, c_hi=c_hi, hi_coff=hc, dw=dw,
                      tap_mask=MASK_A_5x5 if masked else 0, accumulate=True)
    plan2.launch()
    torch.cuda.synchronize()
    assert float((dw * mask - 2 * want).abs().max()) <= 4e-3 * scale
