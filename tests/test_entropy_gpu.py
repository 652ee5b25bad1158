"""GPU parity of the entropy-model kernels against the fixtures generated from the unmodified
reference (tests/golden) and against oracle/ on the same seeded inputs.  Integer results
(symbols, CDF indexes, CDF tables, rANS bytes) are bit-exact; likelihoods are within
LIK_RTOL/LIK_ATOL (fp32 erfc / sigmoid on the device vs on the host)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# A likelihood is a difference of two erfc() values: ulp-level differences between the device and
# host erfc are amplified by cancellation in the tails, so single values agree to ~1e-4 relative
# while the code length sum(-log2 lik) — what bpp is made of — agrees to 1e-5.
LIK_RTOL, LIK_ATOL = 5e-4, 1e-7
BITS_RTOL = 1e-5


def _bits(l):
    return float(-torch.log2(l.double()).sum())


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from masic_b200 import _lib
    _lib.load()                                   # fail loudly if the extension is missing
    return torch.device("cuda:0")


def test_gmm_matches_reference_fixture(dev, golden_dir):
    from masic_b200 import ops
    fx = np.load(golden_dir / "gmm.npz")
    y, sg, mu, w, wl = (_t(fx[k]).to(dev) for k in ("y", "sigma", "mu", "w", "w_logits"))
    y_hat, lik, sym = ops.gmm_likelihood(y, sg, mu, w, K=5, want_symbols=True)
    assert torch.equal(y_hat.cpu(), _t(fx["y_hat"]))                       # bit-exact
    assert torch.equal(sym.cpu(), _t(fx["y_hat"]).to(torch.int32))
    assert torch.allclose(lik.cpu(), _t(fx["lik"]), rtol=LIK_RTOL, atol=LIK_ATOL)
    assert abs(_bits(lik.cpu()) - _bits(_t(fx["lik"]))) <= BITS_RTOL * _bits(_t(fx["lik"]))
    assert float(lik.min()) == pytest.approx(1e-9)
    # fused softmax over K (the engine's path) gives the same likelihoods
    y_hat2, lik2 = ops.gmm_likelihood(y, sg, mu, wl, K=5, weights_are_logits=True)
    assert torch.equal(y_hat2, y_hat)
    assert torch.allclose(lik2.cpu(), _t(fx["lik"]), rtol=LIK_RTOL, atol=LIK_ATOL)


def test_gmm_nhwc_layout_equals_nchw(dev):
    """The engine feeds NHWC buffers and asks for NCHW outputs; both layouts must agree bit for bit."""
    from masic_b200 import _lib, ops
    g = torch.Generator().manual_seed(3)
    n, m, k, h, w = 2, 192, 5, 9, 13
    y = (torch.randn(n, m, h, w, generator=g) * 3).to(dev)
    sg = torch.exp(torch.randn(n, m * k, h, w, generator=g)).to(dev)
    mu = torch.randn(n, m * k, h, w, generator=g).to(dev)
    wl = torch.randn(n, m * k, h, w, generator=g).to(dev)
    y_hat, lik = ops.gmm_likelihood(y, sg, mu, wl, weights_are_logits=True)
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()   # noqa: E731
    y2, l2 = torch.empty_like(y), torch.empty_like(y)
    yq = torch.zeros(n, h, w, 256, dtype=torch.bfloat16, device=dev)
    yn, sn, mn, wn = nhwc(y), nhwc(sg), nhwc(mu), nhwc(wl)
    _lib.check(_lib.load().masic_gmm_likelihood_fwd(yn.data_ptr(), sn.data_ptr(), mn.data_ptr(), wn.data_ptr(), 1, 1,
                                                    n, m, k, h * w, 0.11, y2.data_ptr(), l2.data_ptr(), None, 0,
                                                    yq.data_ptr(), 256, 64, None, 0, 0,
                                                    0, torch.cuda.current_stream().cuda_stream), "gmm")
    assert torch.equal(y2, y_hat) and torch.equal(l2, lik)
    assert torch.equal(yq[..., 64:256].float(), nhwc(y_hat)) and float(yq[..., :64].abs().max()) == 0.0
    # the engine's configuration (NHWC parameters, logits, NCHW outputs, nothing else): the 128-bit-load kernel
    y3, l3 = torch.empty_like(y), torch.empty_like(y)
    _lib.check(_lib.load().masic_gmm_likelihood_fwd(yn.data_ptr(), sn.data_ptr(), mn.data_ptr(), wn.data_ptr(), 1, 1,
                                                    n, m, k, h * w, 0.11, y3.data_ptr(), l3.data_ptr(), None, 0,
                                                    None, 0, 0, None, 0, 0, 0, torch.cuda.current_stream().cuda_stream), "gmm")
    assert torch.equal(y3, y_hat) and torch.equal(l3, lik)


def test_gmm_against_oracle_random(dev):
    from masic_b200 import ops
    from oracle import entropy as E
    g = torch.Generator().manual_seed(11)
    n, m, k, h, w = 1, 192, 5, 19, 34
    y = torch.randn(n, m, h, w, generator=g) * 3
    y.view(-1)[:500] = torch.round(y.view(-1)[:500]) + 0.5                 # exact ties -> round half to even
    sg = torch.exp(torch.rand(n, m * k, h, w, generator=g) * 8 - 3)
    mu = torch.randn(n, m * k, h, w, generator=g) * 2
    wl = torch.randn(n, m * k, h, w, generator=g)
    wt = torch.softmax(wl.view(n, k, m, h, w), dim=1).reshape(n, m * k, h, w)
    y_ref, l_ref = E.gmm_forward(y, sg, mu, wt, k)
    y_hat, lik = ops.gmm_likelihood(y.to(dev), sg.to(dev), mu.to(dev), wt.to(dev))
    assert torch.equal(y_hat.cpu(), y_ref)
    assert torch.allclose(lik.cpu(), l_ref, rtol=LIK_RTOL, atol=LIK_ATOL)
    assert abs(_bits(lik.cpu()) - _bits(l_ref)) <= BITS_RTOL * _bits(l_ref)


def test_gaussian_conditional_indexes_symbols_and_bitstream(dev, golden_dir):
    from masic_b200.entropy_models import GaussianConditional
    fx = np.load(golden_dir / "gc.npz")
    gc = GaussianConditional([float(v) for v in fx["scale_table"]]).eval()
    gc.update()
    gc = gc.to(dev)
    scales, means, y = (_t(fx[k]).to(dev) for k in ("scales", "means", "y"))
    idx = gc.build_indexes(scales)
    assert idx.dtype == torch.int32 and torch.equal(idx.cpu(), _t(fx["indexes"]))          # bit-exact CDF indexes
    y_hat, lik = gc(y, scales, means)
    assert torch.equal(y_hat.cpu(), _t(fx["y_hat"]))
    assert torch.allclose(lik.cpu(), _t(fx["lik"]), rtol=LIK_RTOL, atol=LIK_ATOL)
    assert torch.equal(gc._quantize(y, "symbols", means).cpu(), _t(fx["symbols"]))         # bit-exact symbols
    strings = gc.compress(y, idx, means)
    assert strings[0] == fx["string0"].tobytes()                                           # byte-identical to the reference ext's rANS stream (native coder)
    y_dec = gc.decompress(strings, idx, means)
    assert torch.equal(y_dec.cpu(), _t(fx["y_dec"]))


def test_entropy_bottleneck_forward_tables_and_bitstream(dev, golden_dir):
    from masic_b200.entropy_models import EntropyBottleneck
    fx = np.load(golden_dir / "eb.npz")
    eb = EntropyBottleneck(16).eval()
    sd = {k[3:]: _t(fx[k]) for k in fx.files if k.startswith("sd/")}
    want_cdf = sd["_quantized_cdf"].clone()
    for k in ("_offset", "_quantized_cdf", "_cdf_length"):
        sd[k] = torch.IntTensor()
    eb.load_state_dict(sd)
    eb = eb.to(dev)
    z = _t(fx["z"]).to(dev)
    z_hat, lik = eb(z)
    assert torch.equal(z_hat.cpu(), _t(fx["z_hat"]))                                       # bit-exact
    assert torch.allclose(lik.cpu(), _t(fx["lik"]), rtol=LIK_RTOL, atol=LIK_ATOL)
    assert abs(_bits(lik.cpu()) - _bits(_t(fx["lik"]))) <= BITS_RTOL * _bits(_t(fx["lik"]))
    eb.update()
    assert torch.equal(eb._quantized_cdf.cpu(), want_cdf)                                  # bit-exact tables
    med = eb._medians().detach().view(1, -1, 1, 1)
    assert torch.equal(eb._quantize(z, "symbols", med).cpu(), _t(fx["symbols"]))
    strings = eb.compress(z)
    assert strings[0] == fx["string0"].tobytes()                                           # byte-identical
    z_dec = eb.decompress(strings, z.shape[-2:])
    assert torch.equal(z_dec.cpu(), _t(fx["z_dec"])) and torch.equal(z_dec, z_hat)


def test_quantize_edge_cases(dev):
    from masic_b200 import ops
    x = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 1e9, -3.49999, 0.0, 7.0], device=dev)
    assert ops.quantize(x, None, "symbols").tolist() == [0, 2, 2, 0, -2, 1000000000, -3, 0, 7]
    m = torch.full_like(x, 0.25)
    ref = torch.round(x.cpu() - 0.25) + 0.25
    assert torch.equal(ops.quantize(x, m, "dequantize").cpu(), ref)
    empty = torch.zeros(0, device=dev)
    assert ops.quantize(empty, None, "symbols").numel() == 0


def test_full_size_properties(dev):
    """Config A latents (192 x 76 x 136): likelihoods in (0, 1], integer y_hat, idempotent rounding,
    mixture with identical components == single Gaussian."""
    from masic_b200 import ops
    g = torch.Generator().manual_seed(5)
    n, m, k, h, w = 1, 192, 5, 76, 136
    y = (torch.randn(n, m, h, w, generator=g) * 4).to(dev)
    s1 = torch.exp(torch.randn(n, m, h, w, generator=g)).to(dev)
    m1 = torch.randn(n, m, h, w, generator=g).to(dev)
    sg, mu = s1.repeat(1, k, 1, 1), m1.repeat(1, k, 1, 1)
    wl = torch.randn(n, m * k, h, w, generator=g).to(dev)
    y_hat, lik = ops.gmm_likelihood(y, sg, mu, wl, weights_are_logits=True)
    assert bool(((lik > 0) & (lik <= 1.0 + 1e-6)).all()) and torch.equal(y_hat, torch.round(y_hat))
    y_hat2, _ = ops.gmm_likelihood(y_hat, sg, mu, wl, weights_are_logits=True)
    assert torch.equal(y_hat2, y_hat)
    _, lik_g = ops.gc_likelihood(y_hat, s1, m1)
    # sum_k w_k * p = p  (weights sum to one)
    y_q, lik_gc = ops.gc_likelihood(y, s1, None)
    single = ops.gmm_likelihood(y, sg, torch.zeros_like(mu), wl, weights_are_logits=True)[1]
    assert torch.allclose(single, lik_gc, rtol=1e-4, atol=1e-8)
