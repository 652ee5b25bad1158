"""Host logic of the tensor-core after_conv (masic_b200/convplan.py: fold8_weights_5x5_s1): the Toeplitz-expanded weights
applied to the PIXEL-FOLDED image with a plain 5x5 convolution reproduce ConvTranspose2d / Conv2d (k=5, stride 1, pad 2)
of the reference layer (MASIC.py:600,616; compressai/models/utils.py:137-146) — checked with torch on the CPU in fp64."""
import pytest
import torch
import torch.nn.functional as F


@pytest.mark.parametrize("transposed", [True, False])
@pytest.mark.parametrize("slots", [None, (0, 1, 2, 4, 5, 6)])
def test_folded_weights_reproduce_the_layer(transposed, slots):
    from masic_b200._lib import IMG_XOFF, IMG_XPAD
    from masic_b200.convplan import fold8_weights_5x5_s1
    torch.manual_seed(3)
    c_in, H, W = 6, 11, 40
    w = torch.randn(c_in, 3, 5, 5) if transposed else torch.randn(3, c_in, 5, 5)
    b = torch.randn(3)
    x = torch.randn(2, c_in, H, W)
    ref = (F.conv_transpose2d(x.double(), w.double(), b.double(), stride=1, padding=2) if transposed
           else F.conv2d(x.double(), w.double(), b.double(), stride=1, padding=2))
    big, b48, mask = fold8_weights_5x5_s1(w, b, transposed, slots=slots)
    assert tuple(big.shape) == (48, 64, 5, 5) and tuple(b48.shape) == (48,)
    live = [(t // 5, t % 5) for t in range(25) if (mask >> t) & 1]
    assert live == [(ky, kx) for ky in range(5) for kx in (2, 3)]             # ten live taps: blocks xb and xb + 1
    dead = torch.ones(5, 5, dtype=torch.bool)
    for ky, kx in live:
        dead[ky, kx] = False
    assert float(big[:, :, dead].abs().max()) == 0.0
    # channels-last image with padded rows, pixel x at column x + IMG_XOFF, channel ci at slot slots[ci]
    sl = list(range(c_in)) if slots is None else list(slots)
    img = torch.zeros(2, H, W + IMG_XPAD, 8, dtype=torch.float64)
    for ci, s in enumerate(sl):
        img[:, :, IMG_XOFF:IMG_XOFF + W, s] = x[:, ci].double()
    fold = img.view(2, H, (W + IMG_XPAD) // 8, 64).permute(0, 3, 1, 2)      # one "pixel" = 8 image pixels x 8 channels
    o = F.conv2d(fold, big.double(), b48.double(), stride=1, padding=2)[:, :, :, :W // 8]
    out = torch.stack([o[:, co * 16:co * 16 + 8] for co in range(3)], 1)    # (n, 3, 8, H, W/8): planar blocks
    out = out.permute(0, 1, 3, 4, 2).reshape(2, 3, H, W)
    assert float((out - ref).abs().max()) < 1e-12
