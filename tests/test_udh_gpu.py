"""udh homography front-end on the GPU against oracle/udh.py (pinned to the reference's model.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def test_dlt_inverse_h_adjust_kernel_matches_torch(dev):
    """masic_homography_from_delta vs the reference chain (get_perspective_transform -> inverse -> h_adjust, fp32 torch)
    for displacements up to +-32 px: within 1e-5 relative (the kernel solves in fp64)."""
    from masic_b200 import _lib
    from oracle import udh as OU
    lib = _lib.load()
    g = torch.Generator().manual_seed(1)
    B = 64
    _, _, corners = OU.synthetic_patches(B, seed=5)
    delta = (torch.rand(B, 4, 2, generator=g) - 0.5) * 64
    for img_hw, shift in (((1216, 2176), True), ((512, 512), True), ((256, 256), False)):
        want = OU.homography_from_delta(corners.clone(), delta, img_hw, shift_corners=shift).double()
        out = torch.zeros(B, 3, 3, device=dev)
        c, d = corners.to(dev).contiguous(), delta.to(dev).contiguous()
        _lib.check(lib.masic_homography_from_delta(c.data_ptr(), d.data_ptr(), B, int(shift), img_hw[0], img_hw[1], 256, 256,
                                                   out.data_ptr(), None), "dlt")
        got = out.cpu().double()
        # fp64 ground truth of the same chain bounds the fp32 reference's own error
        truth = OU.homography_from_delta(corners.double(), delta.double(), img_hw, shift_corners=shift)
        scale = truth.abs().amax(dim=(1, 2), keepdim=True)
        err_kernel = ((got - truth).abs() / scale).max()
        err_ref = ((want - truth).abs() / scale).max()
        # the kernel forms corners - corners[0] and corners + delta in fp32 like the reference (that rounding, amplified by
        # the conditioning of the 8x8 system, is all that separates it from the fp64 chain), then solves in fp64
        assert err_kernel <= 1e-5, (err_kernel, err_ref)
        assert ((got - want).abs() / scale).max() <= 1e-5          # "h within 1e-5 of torch"


@pytest.mark.parametrize("gain", [1.0, 200.0])
def test_net_delta_and_homography_match_oracle(dev, gain):
    """The whole front-end: conv_tc plans + max-pools + FC kernels + DLT against the fp32 oracle.  gain scales the last
    Linear so that the displacements are pixels, not the 1e-2 of a random init."""
    from masic_b200.udh import Net
    from oracle import udh as OU
    torch.manual_seed(0)
    oracle = OU.OracleUDHNet(128).eval()
    with torch.no_grad():
        oracle.fc[5].weight.mul_(gain)
        oracle.fc[5].bias.mul_(gain)
    net = Net(patch_size=128).eval()
    net.load_state_dict(oracle.state_dict())
    net = net.to(dev)
    for batch, seed in ((1, 3), (4, 9)):
        a, b, corners = OU.synthetic_patches(batch, seed=seed)
        want = oracle(a, b)
        got = net(a.to(dev), b.to(dev)).cpu()
        assert got.shape == (batch, 4, 2)
        tol = 0.02 * float(want.abs().max()) + 1e-3            # bf16 operands / activations through 8 convs + 2 FCs
        assert float((got - want).abs().max()) <= tol, (float((got - want).abs().max()), tol)
        h_want = OU.homography_from_delta(corners, want, (1216, 2176))
        h_got = net.homography(a.to(dev), b.to(dev), corners.to(dev), (1216, 2176)).cpu()
        # compare where it matters: the displacement of the four image corners under the two homographies (pixels)
        pts = torch.tensor([[0.0, 0.0, 1.0], [2175.0, 0.0, 1.0], [0.0, 1215.0, 1.0], [2175.0, 1215.0, 1.0]]).t()
        pw, pg = h_want @ pts, h_got @ pts
        pw, pg = pw[:, :2] / pw[:, 2:3], pg[:, :2] / pg[:, 2:3]
        disp = float((pw - torch.stack([pts[0], pts[1]])).abs().max())
        assert float((pw - pg).abs().max()) <= 0.02 * disp + 0.05, (float((pw - pg).abs().max()), disp)
        # the DLT of the engine's OWN delta is exact
        # the DLT of the engine's OWN delta is exact (fp64 chain as the yardstick: the fp32 torch chain is itself off by
        # ~3e-4 relative on unshifted corners)
        # (corners + delta is formed in fp32 on both sides, as in the reference: with corner coordinates of ~200 that
        # alone moves the translation entries by ~1e-5 px)
        h_self = OU.homography_from_delta(corners.double(), got.double(), (1216, 2176)).float()
        assert torch.allclose(h_got, h_self, rtol=1e-4, atol=1e-4)
        g = net.get_h(a.to(dev), b.to(dev), corners.to(dev)).cpu()
        g_self = OU.homography_from_delta(corners.double(), got.double(), (256, 256), shift_corners=False).float()
        assert torch.allclose(g, g_self, rtol=1e-4, atol=1e-4)


def test_pair_stream_takes_patches_instead_of_h(dev):
    """images -> criterion without host hops: PairStream.submit_patches runs the udh front-end on the device and feeds
    its h_matrix to the codec engine; identical to calling the net, then submit(h)."""
    from masic_b200.hsic import HSIC
    from masic_b200.udh import Net
    from oracle import udh as OU
    torch.manual_seed(0)
    codec = HSIC().eval().to(dev)
    torch.manual_seed(1)
    udh = Net().eval()
    with torch.no_grad():
        udh.fc[5].weight.mul_(200.0)
    udh = udh.to(dev)
    h, w = 128, 192
    g = torch.Generator().manual_seed(4)
    x1, x2 = torch.rand(1, 3, h, w, generator=g), torch.rand(1, 3, h, w, generator=g)
    a, b, corners = OU.synthetic_patches(1, seed=2)
    Hm = udh.homography(a.to(dev), b.to(dev), corners.to(dev), (h, w))
    ps = codec.pair_stream(h, w, dev, depth=2)
    t0 = ps.submit(x1.pin_memory(), x2.pin_memory(), Hm.cpu().pin_memory())
    r0 = ps.result(t0)
    x2_hat0 = ps.outputs()["x2_hat"].clone()
    ps.attach_homography_net(udh)
    t1 = ps.submit_patches(x1.pin_memory(), x2.pin_memory(), a.pin_memory(), b.pin_memory(), corners.pin_memory())
    r1 = ps.result(t1)
    assert torch.equal(ps.outputs()["x2_hat"], x2_hat0)
    assert r0[:3] == r1[:3]
