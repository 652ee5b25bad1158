"""HSIC.compress / decompress on the GPU path (BASELINE config 4): the per-symbol CDF rule
against a numpy restatement of MASIC.py:1006-1043, and the encode -> decode round trip."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _ref_rows(sig, mu, w, minmax, bound=0.11):
    """numpy float32 restatement of MASIC.py:1006-1043 for one (position, channel): sig/mu/w are (K,)."""
    from scipy.special import erfc
    s = np.arange(0, 2 * minmax + 1, dtype=np.float32)
    pmf = None
    for k in range(sig.shape[0]):
        v = np.abs(s - (mu[k] + np.float32(minmax))).astype(np.float32)
        sc = np.float32(max(sig[k], bound))
        up = (np.float32(0.5) * erfc((np.float32(-(2 ** -0.5)) * ((np.float32(0.5) - v) / sc)).astype(np.float32))).astype(np.float32)
        lo = (np.float32(0.5) * erfc((np.float32(-(2 ** -0.5)) * ((np.float32(-0.5) - v) / sc)).astype(np.float32))).astype(np.float32)
        t = ((up - lo) * w[k]).astype(np.float32)
        pmf = t if pmf is None else (pmf + t).astype(np.float32)
    clip = np.clip(pmf, 1.0 / 65536, 1.0).astype(np.float32)
    q = np.round(clip / np.sum(clip) * 65536)
    return np.concatenate([[0], np.add.accumulate(q)]).astype(np.int64)


def test_symbol_cdf_rows_match_numpy_rule(dev):
    from masic_b200 import _lib
    lib = _lib.load()
    M, K, P, minmax = 16, 5, 7, 9
    g = torch.Generator().manual_seed(3)
    sig = (torch.rand(P, K * M, generator=g) * 3.0)
    sig[0, :8] = 0.01                                   # below the 0.11 bound
    mu = torch.randn(P, K * M, generator=g) * 3.0
    wl = torch.randn(P, K * M, generator=g)
    y = torch.randint(-minmax, minmax + 1, (P, M), generator=g).float()
    ch = torch.tensor([0, 3, 5, 15], dtype=torch.int32)
    d = lambda t: t.to(dev).contiguous()   # noqa: E731
    rows = torch.zeros(P, ch.numel(), 2 * minmax + 2, dtype=torch.int32, device=dev)
    iv = torch.zeros(P, ch.numel(), 3, dtype=torch.int32, device=dev)
    sd, md, wd, yd, cd = d(sig), d(mu), d(wl), d(y), d(ch)
    _lib.check(lib.masic_gmm_symbol_cdfs(sd.data_ptr(), md.data_ptr(), wd.data_ptr(), 1, M, K, P, cd.data_ptr(),
                                         ch.numel(), minmax, 0.11, yd.data_ptr(), rows.data_ptr(), iv.data_ptr(), None),
               "cdfs")
    rows, iv = rows.cpu().numpy(), iv.cpu().numpy()
    wsm = torch.softmax(wl.view(P, K, M), dim=1).numpy()
    worst = 0
    for p in range(P):
        for j, c in enumerate(ch.tolist()):
            ref = _ref_rows(sig.view(P, K, M)[p, :, c].numpy(), mu.view(P, K, M)[p, :, c].numpy(), wsm[p, :, c], minmax)
            worst = max(worst, int(np.abs(rows[p, j] - ref).max()))
            s = int(y[p, c]) + minmax
            assert iv[p, j, 0] == rows[p, j, s] and iv[p, j, 1] == rows[p, j, s + 1] - rows[p, j, s]
            assert iv[p, j, 2] == rows[p, j, -1]
            assert (np.diff(rows[p, j]) >= 1).all()      # every symbol keeps a non-empty interval
    # float32 erfc / exp of two libraries: counts may differ by a unit in the last place of a 16-bit frequency
    assert worst <= 2, worst


@pytest.mark.parametrize("y_order", ["wavefront", "raster"])
def test_compress_decompress_round_trip(dev, tmp_path, y_order):
    from masic_b200.hsic import HSIC
    from oracle.hsic import synthetic_homography
    torch.manual_seed(0)
    net = HSIC().eval()
    with torch.no_grad():
        net.encoder1.g_a_conv4.weight.mul_(8.0)          # non-degenerate latents
        net.encoder2.g_a_conv4.weight.mul_(8.0)
    net = net.to(dev)
    net.update(force=True)
    h, w = 128, 192
    g = torch.Generator().manual_seed(5)
    x1, x2 = torch.rand(1, 3, h, w, generator=g).to(dev), torch.rand(1, 3, h, w, generator=g).to(dev)
    Hm = synthetic_homography(1, seed=1).to(dev)
    with torch.no_grad():
        fwd = net(x1, x2, Hm)
        enc = net.compress(x1, x2, Hm, "pair0", str(tmp_path), y_order=y_order)
    assert torch.equal(enc["y1_hat"], fwd["y1_hat"])
    assert enc["n_symbols"] > 0
    # the range coder lands within a fraction of a percent (+ a few flush bytes) of the ideal code length
    assert enc["y_bytes"] * 8 <= enc["y_bits_ideal"] * 1.002 + 64
    # estimated bits (likelihoods over ALL channels on the unbounded support) bound the coder's ideal length from
    # above: the file skips all-zero channels and renormalises each pmf on [-minmax, minmax] (MASIC.py:925-940,1040)
    est_y = sum(float(torch.log(fwd["likelihoods"][k]).sum()) for k in ("y1", "y2")) / (-math.log(2))
    assert 0.5 * est_y <= enc["y_bits_ideal"] <= 1.01 * est_y + 256
    with torch.no_grad():
        dec = net.decompress(x1, x2, Hm, "pair0", str(tmp_path), device=dev)
    for k in ("y1_hat", "y2_hat", "z1_hat", "z2_hat"):
        assert torch.equal(dec[k], enc[k]), k
    # the decoder reproduces forward()'s reconstructions exactly (same kernels on the same latents)
    assert torch.equal(dec["x1_hat"], fwd["x1_hat"])
    assert torch.equal(dec["x2_hat"], fwd["x2_hat"])
    print(f"{y_order}: enc {enc['enctime'] * 1e3:.1f} ms, dec {dec['dectime'] * 1e3:.1f} ms, {enc['n_symbols']} symbols")


def test_wave_and_raster_orders_cost_the_same_bits(dev, tmp_path):
    """The symbol order changes neither the CDFs nor the ideal code length; the coded sizes differ by flush bytes only."""
    from masic_b200.hsic import HSIC
    from oracle.hsic import synthetic_homography
    torch.manual_seed(0)
    net = HSIC().eval()
    with torch.no_grad():
        net.encoder1.g_a_conv4.weight.mul_(8.0)
        net.encoder2.g_a_conv4.weight.mul_(8.0)
    net = net.to(dev)
    net.update(force=True)
    g = torch.Generator().manual_seed(6)
    x1, x2 = torch.rand(1, 3, 64, 128, generator=g).to(dev), torch.rand(1, 3, 64, 128, generator=g).to(dev)
    Hm = synthetic_homography(1, seed=2).to(dev)
    with torch.no_grad():
        a = net.compress(x1, x2, Hm, "a", str(tmp_path), y_order="wavefront")
        b = net.compress(x1, x2, Hm, "b", str(tmp_path), y_order="raster")
    assert a["n_symbols"] == b["n_symbols"]
    assert abs(a["y_bits_ideal"] - b["y_bits_ideal"]) <= 1e-6 * b["y_bits_ideal"]
    assert abs(a["y_bytes"] - b["y_bytes"]) <= 8


def test_full_size_round_trip_wavefront(dev, tmp_path):
    """BASELINE.json configs[3]: compress / decompress at 1216x2176 (3.75 M coded symbols, 361 waves per view):
    the decoder reproduces forward()'s latents and reconstructions bit for bit and the file lands within 0.1 % of the
    ideal code length of the per-symbol CDFs."""
    from masic_b200.hsic import HSIC
    torch.manual_seed(0)
    net = HSIC().eval()
    with torch.no_grad():
        net.encoder1.g_a_conv4.weight.mul_(8.0)
        net.encoder2.g_a_conv4.weight.mul_(8.0)
    net = net.to(dev)
    net.update(force=True)
    h, w = 1216, 2176
    g = torch.Generator().manual_seed(5)
    x1, x2 = torch.rand(1, 3, h, w, generator=g).to(dev), torch.rand(1, 3, h, w, generator=g).to(dev)
    Hm = torch.tensor([[[1.0, 0.01, 20.0], [0.0, 1.0, 3.0], [1e-6, 0.0, 1.0]]], device=dev)
    with torch.no_grad():
        fwd = net(x1, x2, Hm)
        enc = net.compress(x1, x2, Hm, "full", str(tmp_path))
        dec = net.decompress(x1, x2, Hm, "full", str(tmp_path), device=dev)
    assert enc["n_symbols"] > 1_000_000
    assert enc["y_bytes"] * 8 <= enc["y_bits_ideal"] * 1.001 + 64
    for k in ("y1_hat", "y2_hat", "z1_hat", "z2_hat"):
        assert torch.equal(dec[k], enc[k]), k
    assert torch.equal(dec["y1_hat"], fwd["y1_hat"]) and torch.equal(dec["x1_hat"], fwd["x1_hat"])
    assert torch.equal(dec["x2_hat"], fwd["x2_hat"])
    assert dec["dectime"] < 5.0
