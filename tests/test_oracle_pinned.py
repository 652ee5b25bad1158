"""Pin oracle/ (our CPU restatement) to the fixtures generated from the unmodified reference
(tests/golden/make_golden.py).  CPU-only; integer results must be bit-exact, fp32 results
must be bit-exact too where the oracle runs the same torch-CPU ops in the same order."""
import json
import math

import numpy as np
import pytest
import torch

from oracle import entropy as E
from oracle import hsic as OH
from oracle import refimport


def _t(a):
    return torch.from_numpy(np.asarray(a))


# ----------------------------------------------------------------------------- pmf -> cdf
def test_pmf_to_cdf_kats(golden_dir):
    kat = json.loads((golden_dir / "pmf_cdf_kat.json").read_text())
    assert len(kat["cases"]) >= 40
    for c in kat["cases"]:
        if "error" in c:
            with pytest.raises(ValueError):
                E.pmf_to_quantized_cdf(c["pmf"], kat["precision"])
            with pytest.raises(ValueError):
                E.pmf_to_quantized_cdf_c(c["pmf"], kat["precision"])
            continue
        want = np.asarray(c["cdf"], dtype=np.uint32)
        assert np.array_equal(E.pmf_to_quantized_cdf(c["pmf"], kat["precision"]), want)
        assert np.array_equal(E.pmf_to_quantized_cdf_c(c["pmf"], kat["precision"]), want)


def test_pmf_to_cdf_rejects_bad():
    for bad in ([-0.1, 0.5], [float("nan"), 1.0], [float("inf")], [0.0, 0.0]):
        with pytest.raises(ValueError):
            E.pmf_to_quantized_cdf(bad)
        with pytest.raises(ValueError):
            E.pmf_to_quantized_cdf_c(bad)


def test_pmf_to_cdf_matches_compiled_reference_ext():
    """oracle/_ref/_CXX is the reference's own ops.cpp compiled here; it travels to the GPU box."""
    try:
        cxx = refimport.load_ref_ext("_CXX")
    except ImportError:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(7)
    for _ in range(100):
        p = rng.random(int(rng.integers(1, 80))).astype(np.float32) ** int(rng.integers(1, 10))
        if p.sum() * 65536 < 1:
            continue
        want = np.asarray(cxx.pmf_to_quantized_cdf([float(v) for v in p], 16), dtype=np.uint32)
        assert np.array_equal(E.pmf_to_quantized_cdf_c(p), want)


# ----------------------------------------------------------------------------- EntropyBottleneck
def _eb_params(fx):
    sd = {k[3:]: _t(fx[k]) for k in fx.files if k.startswith("sd/")}
    return E.EBParams.from_state_dict(sd, ""), sd


def test_eb_forward_symbols_tables(golden_dir):
    fx = np.load(golden_dir / "eb.npz")
    p, sd = _eb_params(fx)
    z = _t(fx["z"])
    z_hat, lik = E.eb_forward(p, z)
    assert torch.equal(z_hat, _t(fx["z_hat"]))
    assert torch.equal(lik, _t(fx["lik"]))
    assert torch.equal(E.eb_symbols(p, z), _t(fx["symbols"]))
    off, cdf, ln = E.eb_tables(p)
    assert torch.equal(off, sd["_offset"]) and torch.equal(ln, sd["_cdf_length"])
    assert torch.equal(cdf, sd["_quantized_cdf"])
    off2, cdf2, _ = E.eb_tables(p, use_c=False)
    assert torch.equal(cdf2, cdf) and torch.equal(off2, off)
    assert abs(float(E.eb_aux_loss(p, sd["target"])) - float(fx["aux_loss"])) <= 1e-4 * abs(float(fx["aux_loss"]))
    # decoder side: float(symbol) + median == forward's dequantised value (entropy_models.py:127-134)
    med = p.medians().view(1, -1, 1, 1)
    assert torch.equal(E.eb_symbols(p, z).float() + med, _t(fx["z_dec"]))


def test_eb_init_tables(golden_dir):
    fx = np.load(golden_dir / "eb_init_seed0.npz")
    torch.manual_seed(0)
    m = OH._EBParams(128)
    assert torch.equal(m._biases[0].detach(), _t(fx["biases0"]))      # same RNG consumption as the reference
    off, cdf, ln = E.eb_tables(m.params())
    assert torch.equal(off, _t(fx["offset"])) and torch.equal(ln, _t(fx["length"]))
    assert torch.equal(cdf, _t(fx["cdf"]))
    assert cdf.shape == (128, 23) and int(off[0]) == -10 and cdf[0, :3].tolist() == [0, 1218, 2497]


def test_rans_bitstream_from_oracle_symbols(golden_dir):
    """Symbols/indexes/tables from the oracle -> the reference's compiled rANS coder -> the
    exact bytes the reference produced (entropy_models.py:165-196)."""
    try:
        ans = refimport.load_ref_ext("ans")
    except ImportError:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    fx = np.load(golden_dir / "eb.npz")
    p, _ = _eb_params(fx)
    z = _t(fx["z"])
    off, cdf, ln = E.eb_tables(p)
    sym = E.eb_symbols(p, z)
    idx = E.eb_indexes(z.shape)
    s = ans.RansEncoder().encode_with_indexes(sym[0].reshape(-1).tolist(), idx[0].reshape(-1).tolist(),
                                              cdf.tolist(), ln.tolist(), off.tolist())
    assert bytes(s) == fx["string0"].tobytes()
    dec = ans.RansDecoder().decode_with_indexes(s, idx[0].reshape(-1).tolist(), cdf.tolist(), ln.tolist(),
                                                off.tolist())
    assert dec == sym[0].reshape(-1).tolist()


# ----------------------------------------------------------------------------- GaussianConditional
def test_gc_tables_indexes_likelihood(golden_dir):
    import hashlib
    fx = np.load(golden_dir / "gc.npz")
    table = [float(v) for v in fx["scale_table"]]
    assert np.array_equal(np.asarray(E.default_scale_table(), dtype=np.float32), fx["scale_table"])
    off, cdf, ln = E.gc_tables(table)
    assert torch.equal(off, _t(fx["offset"])) and torch.equal(ln, _t(fx["length"]))
    assert list(cdf.shape) == fx["cdf_shape"].tolist() == [64, 3133]
    assert torch.equal(cdf[fx["cdf_rows"].tolist()], _t(fx["cdf_sel"]))
    assert hashlib.sha256(cdf.numpy().tobytes()).digest() == fx["cdf_sha256"].tobytes()
    scales, means, y = _t(fx["scales"]), _t(fx["means"]), _t(fx["y"])
    assert torch.equal(E.gc_build_indexes(scales, table), _t(fx["indexes"]))
    y_hat, lik = E.gc_forward(y, scales, means)
    assert torch.equal(y_hat, _t(fx["y_hat"])) and torch.equal(lik, _t(fx["lik"]))
    assert torch.equal(E.quantize_symbols(y, means), _t(fx["symbols"]))
    try:
        ans = refimport.load_ref_ext("ans")
    except ImportError:
        return
    idx = E.gc_build_indexes(scales, table)
    s = ans.RansEncoder().encode_with_indexes(E.quantize_symbols(y, means)[0].reshape(-1).tolist(),
                                              idx[0].reshape(-1).tolist(), cdf.tolist(), ln.tolist(), off.tolist())
    assert bytes(s) == fx["string0"].tobytes()      # includes bypass-coded out-of-table symbols


# ----------------------------------------------------------------------------- GMM (what HSIC uses)
def test_gmm_forward(golden_dir):
    fx = np.load(golden_dir / "gmm.npz")
    K = int(fx["K"])
    y, sg, mu, wl = (_t(fx[k]) for k in ("y", "sigma", "mu", "w_logits"))
    n, mk, h, w_ = sg.shape
    w = torch.softmax(wl.view(n, K, mk // K, h, w_), dim=1).reshape(n, mk, h, w_)
    assert torch.equal(w, _t(fx["w"]))
    y_hat, lik = E.gmm_forward(y, sg, mu, w, K)
    assert torch.equal(y_hat, _t(fx["y_hat"]))
    assert torch.equal(lik, _t(fx["lik"]))
    assert float(lik.min()) == pytest.approx(1e-9)       # the floor is exercised


# ----------------------------------------------------------------------------- GDN
def test_gdn(golden_dir):
    fx = np.load(golden_dir / "gdn.npz")
    for tag, inv in (("gdn", False), ("igdn", True)):
        y = OH.gdn(_t(fx[f"{tag}/x"]), _t(fx[f"{tag}/beta"]), _t(fx[f"{tag}/gamma"]), inv)
        assert torch.equal(y, _t(fx[f"{tag}/y"]))


# ----------------------------------------------------------------------------- warp (unpinned by the reference)
def test_warp_regression(golden_dir):
    fx = np.load(golden_dir / "warp.npz")
    img, Hm = _t(fx["img"]), _t(fx["H"])
    assert torch.equal(OH.warp(img, Hm), _t(fx["warped"]))
    m_r, m_l = OH.warp_masks(img, Hm)
    assert torch.equal(m_r, _t(fx["mask_R"])) and torch.equal(m_l, _t(fx["mask_L"]))
    # algebraic cross-check: src_pixel = M^-1 dst_pixel, bilinear, zeros outside
    b, c, h, w = img.shape
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float64), torch.arange(w, dtype=torch.float64), indexing="ij")
    pts = torch.stack((xs, ys, torch.ones_like(xs)), -1).reshape(-1, 3)
    for i in range(b):
        q = pts @ torch.inverse(Hm[i].double()).T
        sx, sy = q[:, 0] / q[:, 2], q[:, 1] / q[:, 2]
        x0, y0 = torch.floor(sx), torch.floor(sy)
        acc = torch.zeros(c, h * w, dtype=torch.float64)
        for dy in (0, 1):
            for dx in (0, 1):
                xi, yi = x0 + dx, y0 + dy
                wgt = (1 - (sx - xi).abs()) * (1 - (sy - yi).abs())
                ok = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
                v = img[i].double()[:, yi.clamp(0, h - 1).long(), xi.clamp(0, w - 1).long()]
                acc += v * (wgt * ok)
        assert (acc.view(c, h, w) - _t(fx["warped"])[i].double()).abs().max() < 1e-4


# ----------------------------------------------------------------------------- HSIC model level
def test_hsic_state_dict_layout(golden_dir):
    lay = json.loads((golden_dir / "hsic_layout.json").read_text())
    torch.manual_seed(0)
    net = OH.OracleHSIC(128, 192, 5)
    sd = net.state_dict()
    got = {k: (list(v.shape), str(v.dtype).replace("torch.", "")) for k, v in sd.items()}
    want = {k: (s, d) for k, s, d in lay["entries"]}
    assert len(want) == 248
    assert got == want
    import hashlib
    for k, h in lay["param_sha256"].items():          # seeded init is bit-identical to the reference's
        assert hashlib.sha256(sd[k].numpy().tobytes()).hexdigest() == h, k


@pytest.mark.parametrize("tag,scale", [("init_128x192", 1.0), ("scaled8_128x192", 8.0), ("scaled50_128x192", 50.0)])
def test_hsic_forward_matches_reference(golden_dir, tag, scale):
    fx = np.load(golden_dir / f"hsic_forward_{tag}.npz")
    h, w = int(fx["h"]), int(fx["w"])
    torch.manual_seed(0)
    net = OH.OracleHSIC(128, 192, 5).eval()
    if scale != 1.0:
        with torch.no_grad():
            net.encoder1.g_a_conv4.weight.mul_(scale)
            net.encoder2.g_a_conv4.weight.mul_(scale)
    g = torch.Generator().manual_seed(int(fx["x_seed"]))
    x1 = torch.rand(1, 3, h, w, generator=g)
    x2 = torch.rand(1, 3, h, w, generator=g)
    Hm = OH.synthetic_homography(1, seed=int(fx["h_seed"]))
    assert torch.equal(Hm, _t(fx["H"]))
    torch.set_num_threads(8)
    out = net(x1, x2, Hm)
    # same torch-CPU kernels in the same order -> bit-identical
    assert torch.equal(out["y1_hat"], _t(fx["y1_hat"]))
    assert torch.equal(out["z1_hat"], _t(fx["z1_hat"]))
    for k in ("y1", "y2", "z1", "z2"):
        assert torch.allclose(out["likelihoods"][k], _t(fx[f"lik_{k}"]), rtol=1e-5, atol=1e-12), k
    cy, cx = h // 2 - 24, w // 2 - 24
    assert torch.allclose(out["x1_hat"][:, :, cy:cy + 48, cx:cx + 48], _t(fx["x1_hat_crop"]), atol=1e-5)
    assert torch.allclose(out["x2_hat"][:, :, cy:cy + 48, cx:cx + 48], _t(fx["x2_hat_crop"]), atol=1e-5)
    assert torch.allclose(out["x1_mask_R"][0, 0, h // 2], _t(fx["mask_R_row"]), atol=1e-6)
    assert torch.allclose(out["x1_mask_L"][0, 0, h // 2], _t(fx["mask_L_row"]), atol=1e-6)
    npx = h * w
    for k in ("y1", "y2", "z1", "z2"):
        bpp = float(torch.log(out["likelihoods"][k]).sum() / (-math.log(2) * npx))
        assert bpp == pytest.approx(float(fx[f"bpp_{k}"]), rel=1e-5, abs=1e-7)
    assert float(torch.mean((out["x1_hat"] - x1) ** 2)) == pytest.approx(float(fx["mse1"]), rel=1e-5)
    assert float(torch.mean((out["x2_hat"] - x2) ** 2)) == pytest.approx(float(fx["mse2"]), rel=1e-5)
    assert int((out["y1_hat"] != 0).sum()) == int(fx["y1_nonzero"])
