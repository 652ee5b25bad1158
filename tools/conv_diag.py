"""GPU diagnostic for the tcgen05 conv kernel: runs a ladder of layer shapes from trivial to
the real MASIC layers, compares against (a) the library's own CUDA-core direct conv and
(b) torch.nn.functional on the same bf16-rounded operands, and prints one line per case.
Never stops at the first failure (first-bring-up tool; the pass/fail tests live in tests/).

    python tools/conv_diag.py [--only NAME] [--big]
"""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from masic_b200 import _lib  # noqa: E402
from masic_b200.convplan import (ACT_LEAKY, ACT_NONE, ACT_RELU, CONV, DECONV_S2, DECONV_S2_SUBPIX,  # noqa: E402
                                 GDN_FWD, GDN_INV, GDN_NONE, MASK_A_5x5, ConvPlan, conv_direct, gdn_prepare)

dev = torch.device("cuda:0")
DT = torch.bfloat16       # 16-bit format of the case being run (run_case(..., dtype=...) sets it): bf16 or fp16


def ref_torch(x_nhwc, c_in, in_coff, w, transposed, k, stride, tap_mask, bias):
    x = x_nhwc[..., in_coff:in_coff + c_in].float().permute(0, 3, 1, 2).contiguous()
    wq = w.to(DT).float()
    if tap_mask:
        m = torch.tensor([(tap_mask >> i) & 1 for i in range(k * k)], device=w.device, dtype=torch.float32).view(1, 1, k, k)
        wq = wq * m
    if transposed:
        y = F.conv_transpose2d(x, wq, bias, stride=stride, padding=k // 2, output_padding=stride - 1)
    else:
        y = F.conv2d(x, wq, bias, stride=stride, padding=k // 2)
    return y.permute(0, 2, 3, 1).contiguous()


def act_ref(y, a):
    if a == ACT_RELU:
        return F.relu(y)
    if a == ACT_LEAKY:
        return F.leaky_relu(y, 0.01)
    return y


def run_case(name, *, kind=CONV, k, stride=1, tap_mask=0, n=1, h, w, c_in, c_out, n_tile, in_cp=None,
             in_coff=0, out_cp=None, out_coff=0, out_fp32=False, act=ACT_NONE, gdn=GDN_NONE, bias=True,
             rowscale=False, transposed=None, seed=0, dtype=torch.bfloat16):
    global DT
    DT = dtype
    name = f"{name}/{'fp16' if dtype == torch.float16 else 'bf16'}"
    torch.manual_seed(seed)
    if kind in (3, 4):
        return run_xfold4(name, n=n, h=h, w=w, c_out=c_out, n_tile=n_tile, gdn=gdn, kind=kind)
    transposed = (kind != CONV) if transposed is None else transposed
    in_cp = in_cp or max(c_in + in_coff, 8)
    x = (torch.randn(n, h, w, in_cp, device=dev) * 1.0).to(DT)
    wshape = (c_in, c_out, k, k) if transposed else (c_out, c_in, k, k)
    wt = torch.randn(*wshape, device=dev) / (c_in * k * k) ** 0.5
    b = torch.randn(c_out, device=dev) * 0.1 if bias else None
    if kind == DECONV_S2:
        ho, wo = 2 * h, 2 * w
    elif kind == CONV and stride == 2:
        ho, wo = h // 2, w // 2
    else:
        ho, wo = h, w
    eff = 4 * c_out if kind == DECONV_S2_SUBPIX else c_out
    c_out_pad = -(-eff // n_tile) * n_tile
    out_cp = out_cp or (c_out_pad + out_coff)
    out = torch.full((n, ho, wo, out_cp), 777.0, device=dev, dtype=torch.float32 if out_fp32 else DT)
    n_nt = c_out_pad // n_tile
    acts = [act] * n_nt if isinstance(act, int) else act
    gb = gg = None
    if gdn:
        gb = torch.sqrt(torch.rand(c_out, device=dev) * 0.5 + 0.75)
        gg = torch.sqrt(torch.rand(c_out, c_out, device=dev) * 0.02 + 0.1 * torch.eye(c_out, device=dev))
    rs = torch.rand(n, ho, wo, 3, device=dev) + 0.5 if rowscale else None
    t0 = time.time()
    try:
        plan = ConvPlan(kind=kind, ksize=k, stride=stride, tap_mask=tap_mask, x=x, in_coff=in_coff, c_in=c_in,
                        weight=wt, transposed=transposed, bias=b, c_out=c_out, n_tile=n_tile, out=out,
                        out_coff=out_coff, act=acts, gdn=gdn, gdn_beta=gb, gdn_gamma=gg, rowscale=rs, rs_off=1)
        plan.launch()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"[{name}] EXCEPTION {type(e).__name__}: {e}", flush=True)
        return False
    # reference
    if kind == DECONV_S2_SUBPIX:
        y = ref_torch(x, c_in, in_coff, wt, True, 5, 2, 0, b)        # (n, 2h, 2w, c_out)
        # kernel layout: [h][w][(py,px,co)]
        y = y.view(n, h, 2, w, 2, c_out).permute(0, 1, 3, 2, 4, 5).reshape(n, h, w, 4 * c_out)
        yd = None
    else:
        y = ref_torch(x, c_in, in_coff, wt, transposed, k, stride if kind == CONV else 2, tap_mask, b)
        yd = conv_direct(x, c_in, wt, transposed=transposed, ksize=k, stride=stride if kind == CONV else 2,
                         tap_mask=tap_mask, bias=b, in_coff=in_coff)
    if gdn:
        beta_p, g32, g16 = gdn_prepare(gb, gg, f16=int(DT == torch.float16))
        norm = torch.einsum("nhwj,ij->nhwi", y * y, g32) + beta_p
        y = y * (torch.rsqrt(norm) if gdn == GDN_FWD else torch.sqrt(norm))
    else:
        parts = [act_ref(y[..., i * n_tile:min((i + 1) * n_tile, eff)], acts[i]) for i in range(n_nt)]
        y = torch.cat([p for p in parts if p.shape[-1] > 0], dim=-1)
    if rs is not None:
        y = y * rs[..., 1:2]
    got = out[..., out_coff:out_coff + eff].float()
    err = (got - y).abs()
    scale = y.abs().max().item() + 1e-9
    maxerr = err.max().item()
    tol = (3e-2 if gdn else 1.2e-2) * scale if not out_fp32 else (3e-2 if gdn else 2e-3) * scale
    if DT == torch.float16:         # 11 significant bits instead of 8: the output rounding and the GDN operands are 8x finer
        tol = ((4e-3 if gdn else 1.5e-3) if not out_fp32 else (4e-3 if gdn else 1e-3)) * scale
    ok = bool(torch.isfinite(got).all()) and maxerr <= tol
    msg = f"[{name}] {'OK ' if ok else 'BAD'} maxerr={maxerr:.4g} scale={scale:.4g} tol={tol:.3g}"
    if yd is not None and not gdn and rs is None:
        yd = torch.cat([act_ref(yd[..., i * n_tile:min((i + 1) * n_tile, eff)], acts[i]) for i in range(n_nt)
                        if i * n_tile < eff], dim=-1)
        msg += f" vs_direct={(got - yd).abs().max().item():.4g}"
    # untouched neighbours (channel padding outside [out_coff, out_coff+c_out_pad) must stay 777)
    if out_coff > 0:
        msg += f" left_pad_ok={bool((out[..., :out_coff].float() == 777).all())}"
    msg += f" work={plan.work_items} smem={plan.smem_bytes} t={time.time() - t0:.2f}s"
    print(msg, flush=True)
    if not ok:
        bad = (err > tol).nonzero()
        print(f"    first bad idx {bad[:5].tolist()} count={bad.shape[0]} of {err.numel()}", flush=True)
        r = err.view(-1, err.shape[-1])
        print(f"    bad rows(pixels)={(r.max(1).values > tol).sum().item()}/{r.shape[0]} "
              f"bad cols(ch)={(r.max(0).values > tol).sum().item()}/{r.shape[1]}", flush=True)
        print(f"    got[0,0,0,:8]={got[0, 0, 0, :8].tolist()}", flush=True)
        print(f"    ref[0,0,0,:8]={y[0, 0, 0, :8].tolist()}", flush=True)
    return ok


def run_xfold4(name, *, n, h, w, c_out, n_tile, gdn, kind=3):
    """g_a_conv1 path: NCHW fp32 image -> padded NHWC bf16 image (pitch 16 / 8) -> XFOLD4 / XFOLD8 plan (+GDN) vs
    torch conv2d."""
    cp = 16 if kind == 3 else 8
    from masic_b200 import _lib as L
    img = torch.rand(n, 3, h, w, device=dev)
    wt = torch.randn(c_out, 3, 5, 5, device=dev) / 75 ** 0.5
    b = torch.randn(c_out, device=dev) * 0.1
    xw = torch.zeros(n, h, w + L.IMG_XPAD, cp, device=dev, dtype=DT)
    L.check(L.load().masic_nchw_to_nhwc_bf16(img.data_ptr(), n, 3, h, w, xw.data_ptr(), cp, w + L.IMG_XPAD, L.IMG_XOFF,
                                             int(DT == torch.float16), torch.cuda.current_stream().cuda_stream), "pack")
    out = torch.zeros(n, h // 2, w // 2, c_out, device=dev, dtype=DT)
    gb = torch.sqrt(torch.rand(c_out, device=dev) * 0.5 + 0.75)
    gg = torch.sqrt(torch.rand(c_out, c_out, device=dev) * 0.02 + 0.1 * torch.eye(c_out, device=dev))
    plan = ConvPlan(kind=kind, ksize=5, stride=2, x=xw, c_in=64, weight=wt, bias=b, c_out=c_out, n_tile=n_tile, out=out,
                    gdn=gdn, gdn_beta=gb, gdn_gamma=gg)
    plan.launch()
    torch.cuda.synchronize()
    y = F.conv2d(img.to(DT).float(), wt.to(DT).float(), b, stride=2, padding=2).permute(0, 2, 3, 1)
    if gdn:
        beta_p, g32, _ = gdn_prepare(gb, gg, f16=int(DT == torch.float16))
        norm = torch.einsum("nhwj,ij->nhwi", y * y, g32) + beta_p
        y = y * torch.rsqrt(norm)
    err = (out.float() - y).abs()
    scale = y.abs().max().item()
    xtol = (4e-3 if DT == torch.float16 else 3e-2) * scale
    ok = bool(torch.isfinite(out.float()).all()) and err.max().item() <= xtol
    print(f"[{name}] {'OK ' if ok else 'BAD'} maxerr={err.max().item():.4g} scale={scale:.4g} work={plan.work_items}", flush=True)
    if not ok:
        bad = (err > xtol).nonzero()
        print("    first bad", bad[:6].tolist(), "count", bad.shape[0], "of", err.numel())
        xs = sorted(set(bad[:, 2].tolist()))
        print("    bad x columns:", xs[:20], " bad y rows:", sorted(set(bad[:, 1].tolist()))[:20])
    return ok


CASES = {
    "1x1_min": dict(k=1, h=16, w=8, c_in=64, c_out=128, n_tile=128, bias=False),
    "1x1_k128_bias_relu_partial": dict(k=1, h=20, w=12, c_in=128, c_out=128, n_tile=128, act=ACT_RELU),
    "1x1_fp32out": dict(k=1, h=16, w=16, c_in=64, c_out=128, n_tile=128, out_fp32=True),
    "3x3_s1": dict(k=3, h=24, w=16, c_in=64, c_out=128, n_tile=128),
    "5x5_s1_192": dict(k=5, h=19, w=34, c_in=192, c_out=128, n_tile=128, act=ACT_RELU),
    "5x5_s2": dict(k=5, stride=2, h=32, w=32, c_in=128, c_out=128, n_tile=128),
    "5x5_s2_gdn": dict(k=5, stride=2, h=64, w=48, c_in=128, c_out=128, n_tile=128, gdn=GDN_FWD),
    "5x5_s2_conv4_fp32": dict(k=5, stride=2, h=38, w=36, c_in=128, c_out=192, n_tile=192, out_fp32=True),
    "5x5_s2_cin16_gdn": dict(k=5, stride=2, h=64, w=64, c_in=16, in_cp=16, c_out=128, n_tile=128, gdn=GDN_FWD),
    "deconv_128": dict(kind=DECONV_S2, k=5, h=19, w=17, c_in=128, c_out=128, n_tile=128),
    "deconv_igdn": dict(kind=DECONV_S2, k=5, h=32, w=24, c_in=192, c_out=128, n_tile=128, gdn=GDN_INV),
    "deconv_192_leaky": dict(kind=DECONV_S2, k=5, h=19, w=34, c_in=128, c_out=192, n_tile=192, act=ACT_LEAKY),
    "deconv_288pad": dict(kind=DECONV_S2, k=5, h=10, w=12, c_in=192, c_out=288, n_tile=192, act=ACT_LEAKY),
    "3x3_cin288": dict(k=3, h=20, w=16, c_in=288, in_cp=384, c_out=384, n_tile=192, rowscale=True),
    "masked5x5": dict(k=5, tap_mask=MASK_A_5x5, h=20, w=24, c_in=192, c_out=384, n_tile=192, out_coff=384, out_cp=768),
    "subpix_deconv4": dict(kind=DECONV_S2_SUBPIX, k=5, h=24, w=16, c_in=128, c_out=3, n_tile=16, out_fp32=True),
    "1x1_gmm_l0": dict(k=1, h=19, w=34, c_in=768, c_out=1152, n_tile=192,
                       act=[ACT_RELU, ACT_RELU, ACT_LEAKY, ACT_LEAKY, ACT_NONE, ACT_RELU]),
    "1x1_deconvk1": dict(k=1, h=16, w=8, c_in=128, c_out=256, n_tile=128, transposed=True),
    "1x1_ntile240": dict(k=1, h=16, w=24, c_in=192, c_out=960, n_tile=192, out_fp32=True, in_coff=64, in_cp=320),
    "xfold4_conv1_gdn": dict(kind=3, k=5, stride=2, n=2, h=48, w=40, c_in=3, c_out=128, n_tile=128, gdn=GDN_FWD),
    "xfold8_conv1_gdn": dict(kind=4, k=5, stride=2, n=2, h=48, w=40, c_in=3, c_out=128, n_tile=128, gdn=GDN_FWD),
    "xfold8_conv1_odd": dict(kind=4, k=5, stride=2, n=1, h=66, w=150, c_in=3, c_out=128, n_tile=128, gdn=GDN_FWD),
    "batch2_s2": dict(k=5, stride=2, n=2, h=32, w=16, c_in=64, c_out=128, n_tile=128),
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    lib = _lib.load()
    print(lib.masic_build_info().decode(), "abi", lib.masic_abi_version(), torch.cuda.get_device_name(0), flush=True)
    res = {}
    for name, kw in CASES.items():
        if a.only and a.only != name:
            continue
        res[name] = run_case(name, **kw)
        res[name + "/fp16"] = run_case(name, dtype=torch.float16, **kw)
    bad = [k for k, v in res.items() if not v]
    print(f"SUMMARY {len(res) - len(bad)}/{len(res)} ok; bad={bad}", flush=True)
    sys.exit(1 if bad else 0)
