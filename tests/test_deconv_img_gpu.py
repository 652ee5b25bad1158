"""g_s_conv4 in col2im form (csrc/deconv_img.cu): ConvTranspose2d(128, 3, k=5, s=2, p=2, op=1) [+ after_gdn] against
torch on the same 16-bit-rounded operands — odd strip counts, several CTA ranges per strip, batch, both formats."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n,h,w", [(1, 8, 14), (1, 16, 20), (2, 32, 45), (1, 64, 136), (1, 304, 544)])
@pytest.mark.parametrize("igdn", [False, True])
def test_deconv_img_matches_torch(n, h, w, dt, igdn):
    from masic_b200.convplan import DeconvImgPlan
    from oracle import hsic as OH
    dev = torch.device("cuda:0")
    torch.manual_seed(h * 131 + w)
    x = torch.randn(n, h, w, 128, device=dev).to(dt)
    wt = torch.randn(128, 3, 5, 5, device=dev) / (128 * 6.25) ** 0.5
    b = torch.randn(3, device=dev) * 0.1
    out = torch.full((n, 3, 2 * h, 2 * w), 777.0, device=dev)
    beta = gamma = None
    if igdn:
        beta = OH.nonneg_init(torch.ones(3) + torch.rand(3)).to(dev)
        gamma = OH.nonneg_init(0.1 * torch.eye(3) + torch.rand(3, 3) * 0.02).to(dev)
    plan = DeconvImgPlan(x=x, weight=wt, bias=b, out=out, igdn_beta=beta, igdn_gamma=gamma)
    plan.launch()
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wt.to(dt).float(), b, stride=2, padding=2, output_padding=1)
    if igdn:
        ref = OH.gdn(ref.cpu(), beta.cpu(), gamma.cpu(), True).to(dev)
    assert bool(torch.isfinite(out).all())
    err = float((out - ref).abs().max())
    assert err <= 2e-4 * float(ref.abs().max()) + 1e-5, err
    plan.launch()                                    # a second launch writes the same image (ring state does not leak)
    torch.cuda.synchronize()
    assert float((out - ref).abs().max()) == err
