"""HSIC.compress / decompress on the GPU path (BASELINE config 4): the per-symbol CDF rule
bit for bit against the reference's own evaluation (torch on CUDA + NumPy, MASIC.py:988-1043) and against the CPU
oracle, and the encode -> decode round trip."""
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


def _reference_rows_torch_cuda(sig, mu, w, minmax, bound, dev):
    """MASIC.py:999-1043 as the reference runs it: the pmf with separate torch ops on the CUDA device (:988-1022;
    `_standardized_cumulative` :738-742, `lower_bound_scale` :1013), the normalisation with NumPy float32 on the host
    (:1040-1043).  sig / mu / w: (P, K, C) tensors (w already soft-maxed) -> int64 rows (P, C, 2*minmax+2)."""
    P, K, C = sig.shape
    L = 2 * minmax + 1
    samples = torch.Tensor(np.arange(0, L)).to(dev).reshape(L, 1, 1).expand(L, P, C)
    sig, w = sig.to(dev), w.to(dev)
    means = mu.to(dev) + minmax
    bnd = torch.Tensor([bound]).to(dev)
    pmf = None
    for k in range(K):
        half = float(0.5)
        values = samples - means[:, k]
        scales = torch.max(sig[:, k], bnd)
        values = abs(values)
        const = float(-(2 ** -0.5))
        upper = half * torch.erfc(const * ((half - values) / scales))
        lower = half * torch.erfc(const * ((-half - values) / scales))
        if pmf is None:
            pmf = (upper - lower) * w[:, k]
        else:
            pmf += (upper - lower) * w[:, k]
    pmf = pmf.cpu().numpy()
    rows = np.zeros((P, C, L + 1), dtype=np.int64)
    for p in range(P):
        for c in range(C):
            clip = np.clip(pmf[:, p, c], 1.0 / 65536, 1.0)
            clip = np.round(clip / np.sum(clip) * 65536)
            rows[p, c, 1:] = [int(v) for v in np.add.accumulate(clip)]
    return rows


@pytest.mark.parametrize("minmax,P,M", [(9, 7, 16), (40, 5, 8), (70, 3, 8), (200, 2, 8)])
def test_symbol_cdf_rows_are_identical_to_the_reference_rule(dev, minmax, P, M):
    """SURVEY a14, bit-exact: `masic_gmm_symbol_cdfs` against (1) the reference's rule evaluated the way the reference
    evaluates it — torch ops on the CUDA device, NumPy float32 normalisation — and (2) oracle.entropy.gmm_symbol_cdf
    (the same rule with torch-CPU erfc).  (1) must agree in EVERY count (row lengths 19 / 81 / 141 / 401 exercise the
    n <= 128 and the split branches of NumPy's pairwise np.sum); (2) may differ where CPU and CUDA erfc differ in the
    last ulp — counted and bounded, never more than one count."""
    from masic_b200 import _lib
    from oracle.entropy import gmm_symbol_cdf
    lib = _lib.load()
    K = 5
    g = torch.Generator().manual_seed(3 + minmax)
    sig = torch.rand(P, K * M, generator=g) * (3.0 if minmax < 50 else 25.0)
    sig[0, :8] = 0.01                                   # below the 0.11 bound
    mu = torch.randn(P, K * M, generator=g) * (3.0 if minmax < 50 else 20.0)
    wl = torch.randn(P, K * M, generator=g)
    y = torch.randint(-minmax, minmax + 1, (P, M), generator=g).float()
    ch = torch.arange(M, dtype=torch.int32)
    d = lambda t: t.to(dev).contiguous()   # noqa: E731
    rows = torch.zeros(P, M, 2 * minmax + 2, dtype=torch.int32, device=dev)
    iv = torch.zeros(P, M, 3, dtype=torch.int32, device=dev)
    sd, md, wd, yd, cd = d(sig), d(mu), d(wl), d(y), d(ch)
    _lib.check(lib.masic_gmm_symbol_cdfs(sd.data_ptr(), md.data_ptr(), wd.data_ptr(), 1, M, K, P, cd.data_ptr(),
                                         M, minmax, 0.11, yd.data_ptr(), rows.data_ptr(), iv.data_ptr(), None), "cdfs")
    rows, iv = rows.cpu().numpy().astype(np.int64), iv.cpu().numpy()
    wsm_cuda = torch.softmax(wd.view(P, K, M), dim=1)                 # what the reference's net hands over (MASIC.py:393)
    ref = _reference_rows_torch_cuda(sig.view(P, K, M), mu.view(P, K, M), wsm_cuda, minmax, 0.11, dev)
    assert np.array_equal(rows, ref), f"{int((rows != ref).sum())} of {rows.size} counts differ from the reference rule"
    # pre-softmaxed weights take the same path
    rows2 = torch.zeros(P, M, 2 * minmax + 2, dtype=torch.int32, device=dev)
    ws = wsm_cuda.reshape(P, K * M).contiguous()
    _lib.check(lib.masic_gmm_symbol_cdfs(sd.data_ptr(), md.data_ptr(), ws.data_ptr(), 0, M, K, P, cd.data_ptr(),
                                         M, minmax, 0.11, None, rows2.data_ptr(), None, None), "cdfs")
    assert np.array_equal(rows2.cpu().numpy(), ref)
    s = (y.long() + minmax).numpy()
    for p in range(P):
        for c in range(M):
            assert iv[p, c, 0] == rows[p, c, s[p, c]] and iv[p, c, 1] == rows[p, c, s[p, c] + 1] - rows[p, c, s[p, c]]
            assert iv[p, c, 2] == rows[p, c, -1]
    assert (np.diff(rows, axis=-1) >= 1).all()           # every symbol keeps a non-empty interval
    # (2) the CPU oracle: identical except where erfc differs between the CPU and CUDA math libraries
    wsm = wsm_cuda.cpu()
    differing = 0
    for p in range(P):
        for c in range(0, M, 3):
            o = gmm_symbol_cdf(sig.view(P, K, M)[p, :, c], mu.view(P, K, M)[p, :, c], wsm[p, :, c], minmax)
            dd = np.abs(np.diff(rows[p, c]) - np.diff(o))
            assert dd.max() <= 1
            differing += int((dd != 0).sum())
    print(f"minmax {minmax}: rows identical to the torch-CUDA rule; {differing} counts differ from the CPU oracle (erfc ulp)")


@pytest.mark.parametrize("y_order", ["wavefront_streams", "wavefront", "raster"])
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
    # the range coder lands within a fraction of a percent (+ a few flush bytes per stream) of the ideal code length
    assert enc["y_bytes"] * 8 <= enc["y_bits_ideal"] * 1.002 + 64 + (80 * 384 if y_order == "wavefront_streams" else 0)
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
    with torch.no_grad():
        c = net.compress(x1, x2, Hm, "c", str(tmp_path), y_order="wavefront_streams")
    assert c["n_symbols"] == a["n_symbols"] and abs(c["y_bits_ideal"] - a["y_bits_ideal"]) <= 1e-6 * a["y_bits_ideal"]
    n_streams = c["n_symbols"] // ((64 // 16) * (128 // 16))        # symbols per position = non-zero channels of both views
    assert 0 <= c["y_bytes"] - a["y_bytes"] <= 12 * n_streams + 16   # per stream: 4-byte length + <= 6 flush bytes


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
    assert enc["y_bytes"] * 8 <= enc["y_bits_ideal"] * 1.001 + 64 + 80 * 384       # 384 streams: table entry + flush bytes each
    for k in ("y1_hat", "y2_hat", "z1_hat", "z2_hat"):
        assert torch.equal(dec[k], enc[k]), k
    assert torch.equal(dec["y1_hat"], fwd["y1_hat"]) and torch.equal(dec["x1_hat"], fwd["x1_hat"])
    assert torch.equal(dec["x2_hat"], fwd["x2_hat"])
    assert dec["dectime"] < 5.0
    # the same file through the host decoder path of the single-stream wave format costs the same bits
    with torch.no_grad():
        enc1 = net.compress(x1, x2, Hm, "full1", str(tmp_path), y_order="wavefront")
        dec1 = net.decompress(x1, x2, Hm, "full1", str(tmp_path), device=dev)
    assert torch.equal(dec1["y2_hat"], dec["y2_hat"]) and torch.equal(dec1["x2_hat"], dec["x2_hat"])
    print(f"1216x2176: device decoder {dec['dectime'] * 1e3:.1f} ms, host decoder {dec1['dectime'] * 1e3:.1f} ms; "
          f"payload {enc['y_bytes']} vs {enc1['y_bytes']} bytes")
