"""Time the tensor-core conv kernel on the real MASIC layer shapes (1216x2176 pair, B=1).

    python tools/conv_perf.py [--iters 5] [--only NAME]

Prints per layer: ms, useful TFLOP/s, fraction of the measured bf16 peak, and the HBM
floor (algorithmic bytes / measured copy bandwidth) for comparison.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from masic_b200.convplan import (ACT_LEAKY, ACT_NONE, ACT_RELU, CONV, DECONV_S2, DECONV_S2_SUBPIX,  # noqa: E402
                                 GDN_FWD, GDN_INV, GDN_NONE, MASK_A_5x5, ConvPlan)

dev = torch.device("cuda:0")
H, W = 1216, 2176

LAYERS = {
    # name: kind, k, stride, h_in, w_in, c_in, c_out, n_tile, gdn, out_fp32, extra
    "g_a_conv1(cin16)": dict(k=5, stride=2, h=H, w=W, c_in=16, c_out=128, n_tile=128, gdn=GDN_FWD),
    "g_a_conv1(xfold4)": dict(kind=3, k=5, stride=2, h=H, w=W, c_in=64, c_out=128, n_tile=128, gdn=GDN_FWD),
    "g_a_conv1(xfold8)": dict(kind=4, k=5, stride=2, h=H, w=W, c_in=64, c_out=128, n_tile=128, gdn=GDN_FWD),
    "g_a_conv2": dict(k=5, stride=2, h=H // 2, w=W // 2, c_in=128, c_out=128, n_tile=128, gdn=GDN_FWD),
    "g_a_conv3": dict(k=5, stride=2, h=H // 4, w=W // 4, c_in=128, c_out=128, n_tile=128, gdn=GDN_FWD),
    "g_a_conv4": dict(k=5, stride=2, h=H // 8, w=W // 8, c_in=128, c_out=192, n_tile=192, out_fp32=True),
    "g_s_conv1": dict(kind=DECONV_S2, k=5, h=H // 16, w=W // 16, c_in=192, c_out=128, n_tile=128, gdn=GDN_INV),
    "g_s_conv2": dict(kind=DECONV_S2, k=5, h=H // 8, w=W // 8, c_in=128, c_out=128, n_tile=128, gdn=GDN_INV),
    "g_s_conv3": dict(kind=DECONV_S2, k=5, h=H // 4, w=W // 4, c_in=128, c_out=128, n_tile=128, gdn=GDN_INV),
    "g_s_conv4(subpix)": dict(kind=DECONV_S2_SUBPIX, k=5, h=H // 2, w=W // 2, c_in=128, c_out=3, n_tile=16, out_fp32=True),
    "h_a_conv1": dict(k=5, stride=1, h=H // 16, w=W // 16, c_in=192, c_out=128, n_tile=128, act=ACT_RELU),
    "ctx_masked": dict(k=5, stride=1, tap_mask=MASK_A_5x5, h=H // 16, w=W // 16, c_in=192, c_out=384, n_tile=192),
    "h_s_conv3x3": dict(k=3, stride=1, h=H // 16, w=W // 16, c_in=288, in_cp=384, c_out=384, n_tile=192),
    "gmm_l0_fused": dict(k=1, h=H // 16, w=W // 16, c_in=768, c_out=3456, n_tile=192, act=ACT_RELU),
    "gmm_l1": dict(k=1, h=H // 16, w=W // 16, c_in=1152, c_out=768, n_tile=192, act=ACT_RELU),
    "gmm_l2": dict(k=1, h=H // 16, w=W // 16, c_in=768, c_out=960, n_tile=192, out_fp32=True),
}
# MMA-shape microbenchmarks (python tools/conv_perf.py --only mb_): same GEMM, different accumulator widths
MICRO = {
    "mb_n64": dict(k=1, h=304, w=272, c_in=512, c_out=512, n_tile=64),
    "mb_n128": dict(k=1, h=304, w=272, c_in=512, c_out=512, n_tile=128),
    "mb_n256": dict(k=1, h=304, w=272, c_in=512, c_out=512, n_tile=256),
    "mb_n16": dict(k=1, h=304, w=272, c_in=512, c_out=16, n_tile=16),
    "mb_n32": dict(k=1, h=304, w=272, c_in=512, c_out=32, n_tile=32),
    "mb_gl0_n128": dict(k=1, h=H // 16, w=W // 16, c_in=768, c_out=3456, n_tile=128, act=ACT_RELU),
    "mb_gl0_n256": dict(k=1, h=H // 16, w=W // 16, c_in=768, c_out=3456, n_tile=256, act=ACT_RELU),
    "mb_gl0_n192": dict(k=1, h=H // 16, w=W // 16, c_in=768, c_out=3456, n_tile=192, act=ACT_RELU),
    "mb_gl0_n128_cg2": dict(k=1, h=H // 16, w=W // 16, c_in=768, c_out=3456, n_tile=128, act=ACT_RELU, cta_pairs=True),
    "mb_gl0_n192_cg2": dict(k=1, h=H // 16, w=W // 16, c_in=768, c_out=3456, n_tile=192, act=ACT_RELU, cta_pairs=True),
    "mb_gl0_n256_cg2": dict(k=1, h=H // 16, w=W // 16, c_in=768, c_out=3456, n_tile=256, act=ACT_RELU, cta_pairs=True),
    "mb_gl1_n192_cg2": dict(k=1, h=H // 16, w=W // 16, c_in=1152, c_out=768, n_tile=192, act=ACT_RELU, cta_pairs=True),
    "mb_gl1_n256_cg2": dict(k=1, h=H // 16, w=W // 16, c_in=1152, c_out=768, n_tile=256, act=ACT_RELU, cta_pairs=True),
    "mb_gl2_n192_cg2": dict(k=1, h=H // 16, w=W // 16, c_in=768, c_out=1920, n_tile=192, out_fp32=True, cta_pairs=True),
    "mb_gl2_n192": dict(k=1, h=H // 16, w=W // 16, c_in=768, c_out=1920, n_tile=192, out_fp32=True),
    "mb_gl2_n128": dict(k=1, h=H // 16, w=W // 16, c_in=768, c_out=1920, n_tile=128, out_fp32=True),
    "mb_ctx_n192": dict(k=5, stride=1, tap_mask=MASK_A_5x5, h=H // 16, w=W // 16, c_in=192, c_out=384, n_tile=192),
    "mb_ctx_n192_cg2": dict(k=5, stride=1, tap_mask=MASK_A_5x5, h=H // 16, w=W // 16, c_in=192, c_out=384, n_tile=192, cta_pairs=True),
    "mb_ctx_n128_cg2": dict(k=5, stride=1, tap_mask=MASK_A_5x5, h=H // 16, w=W // 16, c_in=192, c_out=384, n_tile=128, cta_pairs=True),
    "mb_ctx_n64": dict(k=5, stride=1, tap_mask=MASK_A_5x5, h=H // 16, w=W // 16, c_in=192, c_out=384, n_tile=64),
    "mb_ctx_n96": dict(k=5, stride=1, tap_mask=MASK_A_5x5, h=H // 16, w=W // 16, c_in=192, c_out=384, n_tile=96),
    "mb_hs3_n192": dict(k=3, stride=1, h=H // 16, w=W // 16, c_in=288, in_cp=384, c_out=384, n_tile=192),
    "mb_hs3_n192_cg2": dict(k=3, stride=1, h=H // 16, w=W // 16, c_in=288, in_cp=384, c_out=384, n_tile=192, cta_pairs=True),
    "mb_hs3_n128_cg2": dict(k=3, stride=1, h=H // 16, w=W // 16, c_in=288, in_cp=384, c_out=384, n_tile=128, cta_pairs=True),
    "mb_hs3_n64": dict(k=3, stride=1, h=H // 16, w=W // 16, c_in=288, in_cp=384, c_out=384, n_tile=64),
    "mb_hs3_n96": dict(k=3, stride=1, h=H // 16, w=W // 16, c_in=288, in_cp=384, c_out=384, n_tile=96),
    "mb_ga4_n192": dict(k=5, stride=2, h=H // 8, w=W // 8, c_in=128, c_out=192, n_tile=192, out_fp32=True),
    "mb_ga4_n192_cg2": dict(k=5, stride=2, h=H // 8, w=W // 8, c_in=128, c_out=192, n_tile=192, out_fp32=True, cta_pairs=True),
    "mb_ga4_n48": dict(k=5, stride=2, h=H // 8, w=W // 8, c_in=128, c_out=192, n_tile=48, out_fp32=True),
    "mb_ga4_n32": dict(k=5, stride=2, h=H // 8, w=W // 8, c_in=128, c_out=192, n_tile=32, out_fp32=True),
    "mb_ha1_n128": dict(k=5, stride=1, h=H // 16, w=W // 16, c_in=192, c_out=128, n_tile=128, act=ACT_RELU),
    "mb_ha1_n32": dict(k=5, stride=1, h=H // 16, w=W // 16, c_in=192, c_out=128, n_tile=32, act=ACT_RELU),
    "mb_ha2_n128": dict(k=5, stride=2, h=H // 16, w=W // 16, c_in=128, c_out=128, n_tile=128, act=ACT_RELU),
    "mb_ha2_n32": dict(k=5, stride=2, h=H // 16, w=W // 16, c_in=128, c_out=128, n_tile=32, act=ACT_RELU),
    "mb_ha2_n16": dict(k=5, stride=2, h=H // 16, w=W // 16, c_in=128, c_out=128, n_tile=16, act=ACT_RELU),
    "mb_ha3_n128": dict(k=5, stride=2, h=H // 32, w=W // 32, c_in=128, c_out=128, n_tile=128, out_fp32=True),
    "mb_ha3_n16": dict(k=5, stride=2, h=H // 32, w=W // 32, c_in=128, c_out=128, n_tile=16, out_fp32=True),
    "mb_ha3_n32": dict(k=5, stride=2, h=H // 32, w=W // 32, c_in=128, c_out=128, n_tile=32, out_fp32=True),
    "mb_hs1_n128": dict(kind=DECONV_S2, k=5, h=H // 64, w=W // 64, c_in=128, c_out=128, n_tile=128, act=ACT_LEAKY),
    "mb_hs1_n32": dict(kind=DECONV_S2, k=5, h=H // 64, w=W // 64, c_in=128, c_out=128, n_tile=32, act=ACT_LEAKY),
    "mb_hs2_n192": dict(kind=DECONV_S2, k=5, h=H // 32, w=W // 32, c_in=128, c_out=192, n_tile=192, act=ACT_LEAKY),
    "mb_hs2_n96": dict(kind=DECONV_S2, k=5, h=H // 32, w=W // 32, c_in=128, c_out=192, n_tile=96, act=ACT_LEAKY),
    "mb_hs2_n64": dict(kind=DECONV_S2, k=5, h=H // 32, w=W // 32, c_in=128, c_out=192, n_tile=64, act=ACT_LEAKY),
    "mb_ctx_n128": dict(k=5, stride=1, tap_mask=MASK_A_5x5, h=H // 16, w=W // 16, c_in=192, c_out=384, n_tile=128),
    "mb_hs3_n128": dict(k=3, stride=1, h=H // 16, w=W // 16, c_in=288, in_cp=384, c_out=384, n_tile=128),
    "mb_ha1_n64": dict(k=5, stride=1, h=H // 16, w=W // 16, c_in=192, c_out=128, n_tile=64, act=ACT_RELU),
    "mb_ga4_n96": dict(k=5, stride=2, h=H // 8, w=W // 8, c_in=128, c_out=192, n_tile=96, out_fp32=True),
    "mb_ga4_n64": dict(k=5, stride=2, h=H // 8, w=W // 8, c_in=128, c_out=192, n_tile=64, out_fp32=True),
}


def build(name, kind=CONV, k=1, stride=1, tap_mask=0, h=0, w=0, c_in=0, c_out=0, n_tile=128, gdn=GDN_NONE,
          out_fp32=False, act=ACT_NONE, in_cp=None, cta_pairs=False):
    torch.manual_seed(0)
    in_cp = in_cp or c_in
    if kind in (3, 4):
        x = torch.randn(1, h, w + 8, 16 if kind == 3 else 8, device=dev).to(torch.float16)
        wt = torch.randn(c_out, 3, 5, 5, device=dev) / 75 ** 0.5
    else:
        x = torch.randn(1, h, w, in_cp, device=dev).to(torch.float16)
    transposed = kind not in (CONV, 3, 4)
    if kind not in (3, 4):
        wt = torch.randn(*((c_in, c_out, k, k) if transposed else (c_out, c_in, k, k)), device=dev) / (c_in * k * k) ** 0.5
    b = torch.randn(c_out, device=dev) * 0.1
    ho, wo = (2 * h, 2 * w) if kind == DECONV_S2 else ((h // 2, w // 2) if stride == 2 else (h, w))
    eff = 4 * c_out if kind == DECONV_S2_SUBPIX else c_out
    c_out_pad = -(-eff // n_tile) * n_tile
    out = torch.empty(1, ho, wo, c_out_pad, device=dev, dtype=torch.float32 if out_fp32 else torch.float16)
    gb = gg = None
    if gdn:
        gb = torch.ones(c_out, device=dev)
        gg = torch.sqrt(0.1 * torch.eye(c_out, device=dev) + 1e-3)
    return ConvPlan(kind=kind, ksize=k, stride=stride, tap_mask=tap_mask, x=x, c_in=c_in, weight=wt,
                    transposed=transposed, bias=b, c_out=c_out, n_tile=n_tile, out=out, act=act, gdn=gdn,
                    gdn_beta=gb, gdn_gamma=gg, cta_pairs=cta_pairs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    peaks = {}
    pf = ROOT / "MEASURED_PEAKS.json"
    if pf.exists():
        peaks = json.loads(pf.read_text())
    tf_peak = peaks.get("bf16_tflops", 1590.0)
    bw_peak = peaks.get("hbm_gbs", 6650.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot_ms = 0.0
    layers = dict(LAYERS)
    if a.only and a.only.startswith("mb_"):
        layers = MICRO
    for name, kw in layers.items():
        if a.only and a.only not in name:
            continue
        plan = build(name, **kw)
        for _ in range(2):
            plan.launch()
        torch.cuda.synchronize()
        times = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.launch()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = sorted(times)[len(times) // 2]
        tot_ms += ms
        tf = plan.flops / ms / 1e9
        floor_ms = plan.hbm_bytes / bw_peak / 1e6
        print(f"{name:20s} {ms:8.3f} ms  {tf:8.1f} TF/s ({tf / tf_peak * 100:5.1f}% of {tf_peak:.0f})  "
              f"hbm_floor={floor_ms:.3f} ms  work={plan.work_items} smem={plan.smem_bytes}", flush=True)
        del plan
    print(f"sum {tot_ms:.3f} ms")


if __name__ == "__main__":
    main()
