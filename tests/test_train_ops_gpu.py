"""Kernel-level parity of the training-step kernels (train_entropy.cu, train_elem.cu, train_image.cu) against torch
autograd of the oracle's formulas (oracle/train.py restates the reference; here the same expressions run in fp32 on
the device / CPU and are differentiated by autograd)."""
import math

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


def _close(a, b, rtol, what=""):
    scale = float(b.abs().max()) + 1e-30
    err = float((a.float() - b.float()).abs().max())
    assert err <= rtol * scale, (what, err, scale)


def test_gmm_likelihood_train_matches_autograd(dev):
    from masic_b200 import train_ops as T
    from oracle import train as OT
    torch.manual_seed(1)
    n, M, K, h, w = 2, 192, 5, 6, 7
    y = (torch.randn(n, M, h, w) * 3).requires_grad_()
    noise = torch.rand(n, M, h, w) - 0.5
    sig = (torch.rand(n, M * K, h, w) * 2 - 0.3).clamp_min(0).requires_grad_()      # post-ReLU: zeros, below & above 0.11
    mu = torch.randn(n, M * K, h, w).requires_grad_()
    wl = torch.randn(n, M * K, h, w).requires_grad_()
    wsm = F.softmax(wl.reshape(n, K, M, h, w), dim=1).reshape(n, M * K, h, w)
    out, lik = OT.gmm_forward_train(y, sig, mu, wsm, K, noise)
    c = T.lik_grad_scale(n * h * w * 256)
    (torch.log(lik).sum() * c).backward()
    nh = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().to(dev)   # noqa: E731
    npix = n * h * w
    o_lik = torch.empty(npix, M, device=dev)
    o_yb = torch.empty(npix, M, device=dev, dtype=torch.bfloat16)
    o_dy = torch.empty(npix, M, device=dev)
    o_ds = torch.empty(npix, M * K, device=dev, dtype=torch.bfloat16)
    o_dm, o_dw = torch.empty_like(o_ds), torch.empty_like(o_ds)
    T.gmm_likelihood_train(nh(y), nh(noise), nh(sig), nh(mu), nh(wl), M, K, c, lik=o_lik, y_hat_bf=o_yb, dy=o_dy,
                           dsigma=o_ds, dmu=o_dm, dwl=o_dw)
    torch.cuda.synchronize()
    back = lambda t, ch: t.float().reshape(n, h, w, ch).permute(0, 3, 1, 2).cpu()   # noqa: E731
    assert torch.allclose(back(o_lik, M), lik.detach(), rtol=2e-4, atol=1e-9)
    _close(back(o_yb, M), out.detach(), 4e-3, "y_hat bf16")
    _close(back(o_dy, M), y.grad, 2e-3, "dy")
    # sigma gradient: autograd's relu'(0) = 0 is folded in by the kernel; compare where sigma > 0 and check zeros elsewhere
    ds = back(o_ds, M * K)
    pos = sig.detach() > 0
    _close(ds * pos, sig.grad * pos, 1e-2, "dsigma")
    assert float(ds[~pos].abs().max()) == 0.0
    _close(back(o_dm, M * K), mu.grad, 1e-2, "dmu")
    _close(back(o_dw, M * K), wl.grad, 1e-2, "dwl")


def test_eb_train_and_aux_match_autograd(dev):
    from masic_b200 import train_ops as T
    from oracle import train as OT
    from oracle.hsic import _EBParams
    torch.manual_seed(2)
    Cc, n, h, w = 128, 2, 3, 5
    eb = _EBParams(Cc)
    with torch.no_grad():
        for p in list(eb._matrices) + list(eb._factors):
            p.add_(torch.randn_like(p) * 0.3)
        eb.quantiles.add_(torch.randn_like(eb.quantiles))
    z = (torch.randn(n, Cc, h, w) * 4).requires_grad_()
    noise = torch.rand(n, Cc, h, w) - 0.5
    z_hat, lik = OT.eb_forward_train(OT._EBLive(eb), z, noise)
    c = T.lik_grad_scale(n * h * w * 4096)
    (torch.log(lik).sum() * c).backward()
    mats = [m.detach().to(dev).contiguous() for m in eb._matrices]
    bs = [b.detach().to(dev).contiguous() for b in eb._biases]
    fs = [f.detach().to(dev).contiguous() for f in eb._factors]
    nh = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().to(dev)   # noqa: E731
    o_zh = torch.empty(n, Cc, h, w, device=dev)
    o_lik = torch.empty(n, Cc, h, w, device=dev)
    o_zq = torch.empty(n, h, w, Cc, device=dev, dtype=torch.bfloat16)
    o_dz = torch.empty(n, h, w, Cc, device=dev)
    o_dp = torch.empty(Cc, 58, device=dev)
    T.eb_train(nh(z), nh(noise), n, Cc, h * w, mats, bs, fs, c, z_hat=o_zh, lik=o_lik, zq=o_zq, dz=o_dz, dparams=o_dp)
    torch.cuda.synchronize()
    assert torch.allclose(o_zh.cpu(), z_hat.detach(), atol=1e-6)
    assert torch.allclose(o_lik.cpu(), lik.detach(), rtol=3e-4, atol=1e-9)
    _close(o_dz.permute(0, 3, 1, 2).cpu(), z.grad, 3e-3, "dz")
    dp = o_dp.cpu()
    offs = [0, 3, 12, 21, 30, 33, 36, 39, 42, 45, 46, 49, 52, 55, 58]
    params = list(eb._matrices) + list(eb._biases) + list(eb._factors)
    for i, p in enumerate(params):
        _close(dp[:, offs[i]:offs[i + 1]], p.grad.reshape(Cc, -1), 3e-3, f"param {i}")
    # aux loss
    for p in eb.parameters():
        p.grad = None
    holder = type("H", (), {})()
    total = torch.abs(__import__("oracle.entropy", fromlist=["x"]).eb_logits_cumulative(OT._EBLive(eb, True), eb.quantiles)
                      - eb.target).sum()
    total.backward()
    loss = torch.zeros(1, device=dev)
    dq = torch.empty(Cc, 1, 3, device=dev)
    T.eb_aux_loss(eb.quantiles.detach().to(dev).contiguous(), Cc, mats, bs, fs, eb.target.tolist(), loss, dq)
    torch.cuda.synchronize()
    assert float(loss) == pytest.approx(float(total), rel=1e-5)
    _close(dq.cpu(), eb.quantiles.grad, 1e-4, "dquantiles")


def test_gdn_train_fwd_bwd_pieces(dev):
    """GDN / IGDN as the trainer runs them: x^2 -> 1x1 conv (tensor cores) -> apply; backward a -> 1x1 conv -> b."""
    from masic_b200 import train_ops as T
    from masic_b200.convplan import ConvPlan, WgradPlan, gdn_prepare
    from oracle import train as OT
    from oracle.hsic import _GDNParams
    for inverse in (False, True):
        torch.manual_seed(3)
        Cc, n, h, w = 128, 1, 16, 24
        gp = _GDNParams(Cc, inverse=inverse)
        with torch.no_grad():
            gp.gamma.add_(torch.rand_like(gp.gamma) * 0.05)
            gp.beta.add_(torch.rand_like(gp.beta) * 0.3)
        x0 = torch.randn(n, Cc, h, w)
        xb = x0.permute(0, 2, 3, 1).contiguous().to(dev).to(torch.bfloat16)
        x = xb.float().permute(0, 3, 1, 2).cpu().requires_grad_()
        g0 = torch.randn(n, Cc, h, w)
        gb = g0.permute(0, 2, 3, 1).contiguous().to(dev).to(torch.bfloat16)
        yref = OT.gdn(x, gp)
        (yref * gb.float().permute(0, 3, 1, 2).cpu()).sum().backward()
        beta_p, gamma_p32, _ = gdn_prepare(gp.beta.detach().to(dev), gp.gamma.detach().to(dev))
        sq = torch.empty_like(xb)
        nrm = torch.empty(n, h, w, Cc, device=dev)
        y = torch.empty_like(xb)
        T.gdn_square(xb, sq)
        pn = ConvPlan(ksize=1, stride=1, x=sq, c_in=Cc, weight=gamma_p32.reshape(Cc, Cc, 1, 1), bias=beta_p, c_out=Cc,
                      n_tile=128, out=nrm)
        pn.launch()
        T.gdn_apply(xb, nrm, inverse, y)
        torch.cuda.synchronize()
        _close(y.float().permute(0, 3, 1, 2).cpu(), yref.detach(), 1e-2, "gdn fwd")
        # backward
        t = torch.empty_like(xb)
        dbeta_p = torch.zeros(Cc, device=dev)
        dbias = torch.zeros(Cc, device=dev)
        g = gb.clone()
        T.gdn_bwd_a(g, xb, nrm, inverse, t, dbeta_p)
        v = torch.empty(n, h, w, Cc, device=dev)
        pv = ConvPlan(ksize=1, stride=1, x=t, c_in=Cc, weight=gamma_p32.t().contiguous().reshape(Cc, Cc, 1, 1), c_out=Cc,
                      n_tile=128, out=v)
        pv.launch()
        T.gdn_bwd_b(g, xb, v, dbias)
        dgamma_p = torch.empty(Cc, Cc, device=dev)
        WgradPlan(ksize=1, stride=1, lo=t, c_lo=Cc, hi=sq, c_hi=Cc, dw=dgamma_p).launch()
        dbeta = torch.empty(Cc, device=dev)
        dgamma = torch.empty(Cc, Cc, device=dev)
        T.reparam_bwd(dbeta_p, gp.beta.detach().to(dev), 1e-6, dbeta)
        T.reparam_bwd(dgamma_p, gp.gamma.detach().to(dev).contiguous(), 0.0, dgamma)
        torch.cuda.synchronize()
        _close(g.float().permute(0, 3, 1, 2).cpu(), x.grad, 2e-2, "gdn dx")
        _close(dbias.cpu(), x.grad.sum(dim=(0, 2, 3)), 2e-2, "colsum dx")
        _close(dbeta.cpu(), gp.beta.grad, 2e-2, "dbeta")
        _close(dgamma.cpu(), gp.gamma.grad, 3e-2, "dgamma")


def test_act_bwd_bias_and_latents(dev):
    from masic_b200 import train_ops as T
    torch.manual_seed(4)
    npix, pitch, coff, Cc = 300, 192, 64, 96
    g = torch.randn(npix, pitch, device=dev).to(torch.bfloat16)
    y = torch.randn(npix, 256, device=dev).to(torch.bfloat16)
    y[::7] = 0
    for act, slope in ((1, 0.0), (2, 0.01), (0, 1.0)):
        gg = g.clone()
        bias = torch.zeros(Cc, device=dev)
        T.act_bwd_bias(gg.view(1, 1, npix, pitch), coff, Cc, y.view(1, 1, npix, 256), 128, act, bias)
        yy = y[:, 128:128 + Cc].float()
        want = g[:, coff:coff + Cc].float() * torch.where(yy > 0, torch.ones_like(yy), torch.full_like(yy, slope))
        torch.cuda.synchronize()
        _close(gg[:, coff:coff + Cc], want, 5e-3, f"act {act}")
        assert torch.equal(gg[:, :coff], g[:, :coff]) and torch.equal(gg[:, coff + Cc:], g[:, coff + Cc:])
        _close(bias, gg[:, coff:coff + Cc].float().sum(0), 1e-3, "bias grad")
    # latents
    yl = torch.randn(50, 192, device=dev) * 3
    nz = torch.rand(50, 192, device=dev) - 0.5
    ya = torch.empty(50, 192, device=dev, dtype=torch.bfloat16)
    yn = torch.empty(50, 192, device=dev, dtype=torch.bfloat16)
    T.latent_prep_train(yl, nz, ya, yn)
    _close(ya, yl.abs(), 4e-3)
    _close(yn, yl + nz, 4e-3)
    d1, d2, d3 = (torch.randn(50, 192, device=dev).to(torch.bfloat16) for _ in range(3))
    dl = torch.randn(50, 192, device=dev)
    out = torch.empty(50, 192, device=dev, dtype=torch.bfloat16)
    T.latent_merge_bwd(yl, dl, d1, d2, d3, out)
    _close(out, dl + d1.float() + d2.float() + torch.sign(yl) * d3.float(), 4e-3)
    # mask fuse
    npx, c2, m = 40, 384, 192
    P2, C2 = (torch.randn(npx, c2, device=dev).to(torch.bfloat16) for _ in range(2))
    y1w, nz2 = torch.randn(npx, m, device=dev), torch.rand(npx, m, device=dev) - .5
    mw = torch.softmax(torch.randn(npx, 3, device=dev), dim=1)
    fused = torch.empty(npx, 2 * c2 + m, device=dev, dtype=torch.bfloat16)
    T.mask_fuse_fwd(P2, C2, y1w, nz2, mw, fused)
    want = torch.cat((P2.float() * mw[:, 0:1], C2.float() * mw[:, 1:2], (y1w + nz2) * mw[:, 2:3]), dim=1)
    _close(fused, want, 4e-3)
    gf = torch.randn(npx, 2 * c2 + m, device=dev).to(torch.bfloat16)
    dP, dC = torch.empty_like(P2), torch.empty_like(C2)
    dY = torch.empty(npx, m, device=dev, dtype=torch.bfloat16)
    dmw = torch.empty(npx, 3, device=dev)
    T.mask_fuse_bwd(gf, P2, C2, y1w, nz2, mw, dP, dC, dY, dmw)
    gff = gf.float()
    _close(dP, gff[:, :c2] * mw[:, 0:1], 4e-3)
    _close(dC, gff[:, c2:2 * c2] * mw[:, 1:2], 4e-3)
    _close(dY, gff[:, 2 * c2:] * mw[:, 2:3], 4e-3)
    want_mw = torch.stack(((gff[:, :c2] * P2.float()).sum(1), (gff[:, c2:2 * c2] * C2.float()).sum(1),
                           (gff[:, 2 * c2:] * (y1w + nz2)).sum(1)), dim=1)
    _close(dmw, want_mw, 1e-4)


def test_image_domain_backward(dev):
    from masic_b200 import _lib, train_ops as T
    from masic_b200._lib import check
    from oracle import hsic as OH
    from oracle import train as OT
    lib = _lib.load()
    torch.manual_seed(5)
    n, h, w = 2, 24, 40
    # ---- warp backward
    src = torch.rand(n, 3, h, w, requires_grad=True)
    Hm = OH.synthetic_homography(n, seed=3)
    Hm[:, 0, 2] *= 0.2
    out = OH.warp(src, Hm)
    g = torch.randn(n, 3, h, w)
    (out * g).sum().backward()
    Tm = torch.empty(n, 3, 3, device=dev, dtype=torch.float64)
    check(lib.masic_warp_prepare(Hm.to(dev).data_ptr(), n, h, w, h, w, 0, Tm.data_ptr(), None), "prep")
    ds = torch.zeros(n, 3, h, w, device=dev)
    half = (g / 2).to(dev)
    T.warp_bwd(half, half, Tm, ds)
    torch.cuda.synchronize()
    _close(ds.cpu(), src.grad, 2e-3, "warp bwd")
    # ---- small convs: conv s2 3x3 + relu, conv s1 5x5 (6->3), transposed s1 5x5 (6->3)
    for (tr, k, s, c0, c1, co, relu) in ((0, 3, 2, 3, 0, 6, True), (0, 5, 1, 3, 3, 3, False), (1, 5, 1, 3, 3, 3, False),
                                         (0, 3, 2, 1, 0, 3, True)):
        a = torch.randn(n, c0, h, w, requires_grad=True)
        b = torch.randn(n, c1, h, w, requires_grad=True) if c1 else None
        wt = (torch.randn(c0 + c1, co, k, k) if tr else torch.randn(co, c0 + c1, k, k)).requires_grad_()
        bias = torch.randn(co, requires_grad=True)
        x = torch.cat((a, b), dim=1) if c1 else a
        o = F.conv_transpose2d(x, wt, bias, stride=1, padding=k // 2) if tr else F.conv2d(x, wt, bias, stride=s, padding=k // 2)
        oa = F.relu(o) if relu else o
        go = torch.randn_like(oa)
        (oa * go).sum().backward()
        da = torch.empty(n, c0, h, w, device=dev)
        db_ = torch.empty(n, c1, h, w, device=dev) if c1 else None
        dw = torch.zeros_like(wt, device=dev)
        dbias = torch.zeros(co, device=dev)
        T.conv_small_bwd(a.detach().to(dev), None if b is None else b.detach().to(dev), wt.detach().to(dev), tr, co, k, s,
                         go.to(dev), act_out=oa.detach().to(dev) if relu else None, din0=da, din1=db_, dweight=dw,
                         dbias=dbias)
        torch.cuda.synchronize()
        _close(da.cpu(), a.grad, 1e-4, "small dgrad a")
        if c1:
            _close(db_.cpu(), b.grad, 1e-4, "small dgrad b")
        _close(dw.cpu(), wt.grad, 1e-4, "small wgrad")
        _close(dbias.cpu(), bias.grad, 1e-4, "small bias")
    # ---- GDN(3) backward both directions
    for inverse in (False, True):
        gp = OH._GDNParams(3, inverse=inverse)
        with torch.no_grad():
            gp.gamma.add_(torch.rand(3, 3) * 0.1)
        x = torch.randn(n, 3, h, w, requires_grad=True)
        go = torch.randn(n, 3, h, w)
        (OT.gdn(x, gp) * go).sum().backward()
        dx = torch.empty(n, 3, h, w, device=dev)
        dbp, dgp = torch.zeros(3, device=dev), torch.zeros(3, 3, device=dev)
        bd, gd = gp.beta.detach().to(dev), gp.gamma.detach().to(dev).contiguous()
        T.gdn_small_bwd(x.detach().to(dev), go.to(dev), bd, gd, inverse, dx, dbp, dgp)
        dbeta, dgamma = torch.empty(3, device=dev), torch.empty(3, 3, device=dev)
        T.reparam_bwd(dbp, bd, 1e-6, dbeta)
        T.reparam_bwd(dgp, gd, 0.0, dgamma)
        torch.cuda.synchronize()
        _close(dx.cpu(), x.grad, 1e-4, "gdn3 dx")
        _close(dbeta.cpu(), gp.beta.grad, 1e-3, "gdn3 dbeta")
        _close(dgamma.cpu(), gp.gamma.grad, 1e-3, "gdn3 dgamma")
    # ---- softmax backward, mse grad, colsum, wgrad_small
    lg = torch.randn(n, 3, 5, 7, requires_grad=True)
    sm = F.softmax(lg, dim=1)
    gsm = torch.randn(n, 3, 5, 7)
    (sm * gsm).sum().backward()
    nh = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().to(dev)   # noqa: E731
    dl = torch.empty(n, 3, 5, 7, device=dev)
    T.softmax_channels_bwd(nh(sm), nh(gsm), n, 3, 35, dl)
    _close(dl.cpu(), lg.grad, 1e-5, "softmax bwd")
    xh, xx, ad = torch.rand(n, 3, h, w, device=dev), torch.rand(n, 3, h, w, device=dev), torch.randn(n, 3, h, w, device=dev)
    gm = torch.empty_like(xh)
    T.mse_grad(xh, xx, 0.37, gm, addend=ad)
    _close(gm, 0.37 * (xh - xx) + ad, 1e-6)
    cs = torch.zeros(3, device=dev)
    T.colsum_nchw(ad, cs)
    _close(cs, ad.sum(dim=(0, 2, 3)), 1e-4)
    lo = torch.randn(n, h // 2, w // 2, 128, device=dev).to(torch.bfloat16)
    hi = torch.randn(n, 3, h, w, device=dev)
    wz = torch.zeros(128, 3, 5, 5, device=dev, requires_grad=True)
    o = F.conv2d(hi, wz, stride=2, padding=2)
    (o * lo.float().permute(0, 3, 1, 2)).sum().backward()
    dws = torch.zeros(128, 3, 5, 5, device=dev)
    T.wgrad_small(lo, 128, hi, dws)
    torch.cuda.synchronize()
    _close(dws, wz.grad, 1e-4, "wgrad small")


def test_pack_batch_equals_the_per_layer_pack(dev):
    """masic_pack_batch_* (every re-pack of a training step in one launch; plain conv / transposed-conv layers through a
    shared-memory tile, the special kinds element-wise) writes exactly what masic_pack_conv_weights writes — every layer
    shape class of the trainer: 5x5 / 3x3 / 1x1, Conv2d and ConvTranspose2d layouts, stride-1 transposed (flipped) packs,
    channel counts that are not multiples of 64 or of the n-tile, the XFOLD8 and sub-pixel kinds, padded biases."""
    from masic_b200 import _lib
    from masic_b200.convplan import PackBatch, PackedConv
    g = torch.Generator().manual_seed(3)
    CONV, DEC = _lib.CONV, _lib.DECONV_S2
    cases = [  # kind, k, c_in, c_out, n_tile, transposed
        (CONV, 5, 128, 128, 128, False), (CONV, 5, 128, 192, 192, False), (CONV, 1, 768, 3456, 192, False),
        (CONV, 3, 288, 384, 192, False), (CONV, 5, 192, 384, 192, False), (CONV, 1, 1152, 768, 192, True),
        (CONV, 3, 288, 384, 192, True), (CONV, 5, 128, 128, 128, True), (DEC, 5, 128, 128, 128, True),
        (DEC, 5, 192, 128, 128, True), (DEC, 5, 128, 128, 128, False), (DEC, 5, 320, 192, 192, False),
        (CONV, 5, 48, 80, 16, False), (_lib.CONV_XFOLD8, 5, 64, 128, 128, False), (_lib.DECONV_S2_SUBPIX, 5, 128, 3, 16, True),
    ]
    jobs, want = [], []
    for kind, k, ci, co, nt, tr in cases:
        if kind == _lib.CONV_XFOLD8:
            w = torch.randn(co, 3, k, k, generator=g).to(dev)
        elif tr:
            w = torch.randn(ci, co, k, k, generator=g).to(dev)
        else:
            w = torch.randn(co, ci, k, k, generator=g).to(dev)
        b = torch.randn(co, generator=g).to(dev)
        ref = PackedConv(kind=kind, ksize=k, c_in=ci, c_out=co, n_tile=nt, weight=w, transposed=tr, bias=b)
        pk = PackedConv(kind=kind, ksize=k, c_in=ci, c_out=co, n_tile=nt, weight=torch.zeros_like(w), transposed=tr,
                        bias=torch.zeros_like(b))
        pk.w_packed.fill_(7.0)                       # stale contents must be overwritten, padding included
        jobs.append((pk, w, b))
        want.append(ref)
    PackBatch(jobs).launch()
    torch.cuda.synchronize()
    for (pk, _, _), ref, case in zip(jobs, want, cases):
        assert torch.equal(pk.w_packed.view(torch.int16), ref.w_packed.view(torch.int16)), case
        assert torch.equal(pk.bias, ref.bias), case
