"""GPU parity of the tcgen05 implicit-GEMM conv kernel: every layer shape class MASIC uses
(strided 5x5, transposed 5x5 in 4-phase and sub-pixel form, masked 5x5, 3x3, 1x1 incl. the
k=1 ConvTranspose2d, fused GDN / IGDN, per-pixel scaling, channel-slice concatenation, batch)
against torch fp32 convolutions on the same bf16-rounded operands and against the library's
CUDA-core direct conv.  Tolerance: bf16 output rounding (2^-8 relative to the tensor's range)."""
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))


@pytest.fixture(scope="module")
def diag():
    assert torch.cuda.is_available()
    import conv_diag
    return conv_diag


def _names():
    import importlib.util
    spec = importlib.util.spec_from_file_location("conv_diag_names", Path(__file__).resolve().parents[1] / "tools" / "conv_diag.py")
    # only the CASES dict keys are needed at collection time; avoid touching CUDA here
    src = Path(spec.origin).read_text()
    import re
    return re.findall(r'^    "([a-z0-9_]+)": dict\(', src, flags=re.M)


@pytest.mark.parametrize("fmt", ["bf16", "fp16"])      # the training step's format / the inference engines' format
@pytest.mark.parametrize("name", _names())
def test_conv_case(diag, name, fmt):
    assert diag.run_case(name, dtype=torch.float16 if fmt == "fp16" else torch.bfloat16, **diag.CASES[name])


def test_plan_rejects_unsupported(diag):
    from masic_b200 import _lib
    from masic_b200.convplan import ConvPlan, GDN_FWD
    dev = torch.device("cuda:0")
    x = torch.zeros(1, 16, 16, 64, dtype=torch.bfloat16, device=dev)
    out = torch.zeros(1, 16, 16, 64, dtype=torch.bfloat16, device=dev)
    w = torch.zeros(64, 64, 1, 1, device=dev)
    with pytest.raises(_lib.MasicError):        # fused GDN needs c_out == n_tile == 128
        ConvPlan(ksize=1, x=x, c_in=64, weight=w, c_out=64, n_tile=64, out=out, gdn=GDN_FWD,
                 gdn_beta=torch.ones(64, device=dev), gdn_gamma=torch.eye(64, device=dev))
    with pytest.raises(_lib.MasicError):        # n_tile must divide into 16s
        ConvPlan(ksize=1, x=x, c_in=64, weight=w, c_out=64, n_tile=24, out=out)


def _bf(t, dt=torch.bfloat16):
    return t.to(dt).float()


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("c,h,w", [(32, 40, 24), (64, 33, 50), (96, 48, 40)])
def test_residual_epilogue(diag, c, h, w, dt):
    """ResidualBlock / Enhancement_Block tail (layers.py:175-190, MASIC.py:156-164) fused into the conv epilogue:
    out = leaky_relu(conv3x3(x) + b) + res0 (+ res1 at a channel offset of a wider buffer)."""
    import torch.nn.functional as F
    from masic_b200.convplan import ACT_LEAKY, ConvPlan
    dev = torch.device("cuda:0")
    torch.manual_seed(c)
    x = torch.randn(2, h, w, c, device=dev).to(dt)
    r0 = torch.randn(2, h, w, c, device=dev).to(dt)
    r1 = torch.randn(2, h, w, c + 32, device=dev).to(dt)
    wt = torch.randn(c, c, 3, 3, device=dev) / (9 * c) ** 0.5
    b = torch.randn(c, device=dev) * 0.1
    ref = F.leaky_relu(F.conv2d(x.float().permute(0, 3, 1, 2), _bf(wt, dt), b, padding=1), 0.01).permute(0, 2, 3, 1)
    for two in (False, True):
        out = torch.zeros(2, h, w, c + 16, dtype=dt, device=dev)
        ConvPlan(ksize=3, x=x, c_in=c, weight=wt, bias=b, c_out=c, n_tile=c, out=out, out_coff=16, act=ACT_LEAKY,
                 residual0=r0, residual1=r1 if two else None, res1_coff=32).launch()
        torch.cuda.synchronize()
        want = ref + r0.float() + (r1[..., 32:32 + c].float() if two else 0.0)
        got = out[..., 16:16 + c].float()
        tol = 2.0 ** (-8 if dt == torch.bfloat16 else -11) * float(want.abs().max()) + 1e-3
        assert float((got - want).abs().max()) <= tol
        assert float(out[..., :16].abs().max()) == 0.0            # the channel slice before out_coff is untouched


def test_grouped_launch(diag):
    """Three 1x1 layers with a common K run as one block-diagonal plan (the GMM parameter branches, MASIC.py:338-376):
    per n-tile input-channel offset, output channel and output image."""
    import torch.nn.functional as F
    from masic_b200.convplan import ACT_LEAKY, ACT_NONE, ACT_RELU, ConvPlan, PackedConv
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    B, h, w, K = 2, 19, 34, 128
    x = torch.randn(B, h, w, 3 * K, device=dev).to(torch.bfloat16)
    ws = [torch.randn(co, K, 1, 1, device=dev) / K ** 0.5 for co in (192, 192, 384)]
    bs = [torch.randn(co, device=dev) * 0.1 for co in (192, 192, 384)]
    acts = (ACT_RELU, ACT_NONE, ACT_LEAKY)
    pc = PackedConv(ksize=1, c_in=K, c_out=768, n_tile=192, weight=torch.cat(ws, 0), bias=torch.cat(bs))
    # branch 0 -> image block 0 channels 0..191, branch 1 -> image block 1 channels 0..191, branch 2 -> block 2 (384 ch)
    out = torch.zeros(3 * B, h, w, 384, dtype=torch.float32, device=dev)
    ConvPlan(packed=pc, x=x, out=out, act=[acts[0], acts[1], acts[2], acts[2]],
             nt_in_coff=[0, K, 2 * K, 2 * K], nt_out_coff=[0, 0, 0, 192], nt_out_img=[0, B, 2 * B, 2 * B]).launch()
    torch.cuda.synchronize()
    for g in range(3):
        xi = x[..., g * K:(g + 1) * K].float().permute(0, 3, 1, 2)
        y = F.conv2d(xi, _bf(ws[g]), bs[g])
        y = F.relu(y) if acts[g] == ACT_RELU else (F.leaky_relu(y, 0.01) if acts[g] == ACT_LEAKY else y)
        got = out[g * B:(g + 1) * B, :, :, :ws[g].shape[0]]
        assert torch.allclose(got, y.permute(0, 2, 3, 1), atol=2e-3, rtol=1e-3), g
    assert float(out[:2 * B, :, :, 192:].abs().max()) == 0.0
