"""HSIC.compress / HSIC.decompress on the sm_100a kernels (reference: coremasic/mywork/MASIC.py:855-1158, :1161-1408).

What the reference does, and what is kept:
  * side information: `EntropyBottleneck.compress` of z1, z2 through the reference's rANS extension, behind the
    same file header (`<name>.npz`: uint16 H, W | uint16 len(z1), minmax1 | M/8-byte non-zero-channel bitmap | z1 |
    uint16 len(z2), minmax2 | bitmap | z2 — MASIC.py:922-952).  Byte-identical given identical z symbols.
  * y1, y2: one range-coded symbol per (row, column, non-zero channel) in raster order, each under its own
    integer CDF built from the K=5 Gaussian-mixture parameters by the rule of MASIC.py:1006-1043.  The encoder
    knows y_hat, so ALL parameters come from one forward pass of the engine and ALL coding intervals from one
    kernel (`masic_gmm_symbol_cdfs`) — the reference recomputes the context model and the parameter nets per
    pixel in Python.  The decoder is autoregressive (the mask-A context conv needs the already-decoded
    neighbourhood) and walks the pixels in order, evaluating the SAME conv kernels on a 5x5 crop / a single
    pixel so that it reproduces the encoder's parameters bit for bit.
  * the range coder itself: the reference calls the PyPI package `range_coder`, which it neither vendors nor
    pins and which is not installable here; `masic_range_encode` / `masic_range_decode_rows` (csrc/cdf.cu) are a
    plain 32-bit range coder with the same interface.  `<name>.bin` therefore has this library's byte format.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import MasicError, check
from .convplan import ACT_LEAKY, ACT_NONE, ACT_RELU, MASK_A_5x5, ConvPlan
from .engine import SCALE_BOUND, HSICEngine


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _nonzero_channels(y_hat_nchw: torch.Tensor):
    """MASIC.py:925-940: bitmap of channels with any non-zero symbol, and minmax = max(|y_hat|, 1)."""
    a = y_hat_nchw.abs()
    flag = (a.sum(dim=(0, 2, 3)) > 0).to(torch.uint8).cpu().numpy().astype(np.int64)
    minmax = int(max(float(a.max()), 1.0))
    return flag, minmax


def _pack_flags(flag: np.ndarray) -> bytes:
    # np.packbits(np.reshape(flag, [8, M // 8])) flattens first: identical to packing the flat array (MASIC.py:932)
    return np.packbits(np.reshape(flag, [8, flag.shape[0] // 8])).astype(np.uint8).tobytes()


class _PixelModel:
    """GMM parameters of ONE latent position through the engine's own conv kernels: the mask-A context conv on a
    5x5 crop of the decoded latents, then the three 1x1 parameter branches on a 1-pixel image."""

    def __init__(self, eng: HSICEngine, tag: str):
        M, K, dev = eng.M, eng.K, eng.dev
        self.eng, self.tag, self.M, self.K = eng, tag, M, K
        self.right = tag == "R"
        cin = 5 * M if self.right else 4 * M
        MK = M * K
        bf, f32 = torch.bfloat16, torch.float32
        z = lambda *s, dtype=bf: torch.zeros(*s, dtype=dtype, device=dev)   # noqa: E731
        self.crop = z(1, 5, 5, M)
        self.ctx_out = z(1, 5, 5, cin)
        self.rs = z(1, 5, 5, 3, dtype=f32) if self.right else None
        self.px_in = z(1, 1, 1, cin)
        self.l0, self.l1, self.l1w = z(1, 1, 1, 18 * M), z(1, 1, 1, 8 * M), z(1, 1, 1, MK)
        self.sig, self.mu, self.wl = z(1, 1, 1, MK, dtype=f32), z(1, 1, 1, MK, dtype=f32), z(1, 1, 1, MK, dtype=f32)
        pk = eng.packs
        # exactly the epilogue configuration of HSICEngine._build / _gmm_net
        self.plans = [
            ConvPlan(packed=pk[f"{tag}.context"], stride=1, tap_mask=MASK_A_5x5, x=self.crop, out=self.ctx_out,
                     out_coff=2 * M, rowscale=self.rs, rs_off=1)]
        self.tail = [
            ConvPlan(packed=pk[f"{tag}.gmm.l0"], x=self.px_in, out=self.l0, act=[ACT_RELU] * 6 + [ACT_LEAKY] * 12),
            ConvPlan(packed=pk[f"{tag}.gmm.sigma.l1"], x=self.l0, out=self.l1, in_coff=0, out_coff=0, act=ACT_RELU),
            ConvPlan(packed=pk[f"{tag}.gmm.means.l1"], x=self.l0, out=self.l1, in_coff=6 * M, out_coff=4 * M, act=ACT_LEAKY),
            ConvPlan(packed=pk[f"{tag}.gmm.weights.l1"], x=self.l0, out=self.l1w, in_coff=12 * M, act=ACT_LEAKY),
            ConvPlan(packed=pk[f"{tag}.gmm.sigma.l2"], x=self.l1, out=self.sig, in_coff=0, act=ACT_RELU),
            ConvPlan(packed=pk[f"{tag}.gmm.means.l2"], x=self.l1, out=self.mu, in_coff=4 * M),
            ConvPlan(packed=pk[f"{tag}.gmm.weights.l2"], x=self.l1w, out=self.wl),
        ]

    def params_at(self, ypad: torch.Tensor, gmm_in: torch.Tensor, h: int, w: int):
        M = self.M
        self.crop.copy_(ypad[:, h:h + 5, w:w + 5, :])
        if self.right:
            self.rs.copy_(self.eng.mask_weights[:, h:h + 1, w:w + 1, :].expand(1, 5, 5, 3))
        self.plans[0].launch()
        self.px_in[..., :2 * M].copy_(gmm_in[:, h:h + 1, w:w + 1, :2 * M])
        self.px_in[..., 2 * M:4 * M].copy_(self.ctx_out[:, 2:3, 2:3, 2 * M:4 * M])
        if self.right:
            self.px_in[..., 4 * M:].copy_(gmm_in[:, h:h + 1, w:w + 1, 4 * M:])
        for p in self.tail:
            p.launch()
        return self.sig, self.mu, self.wl


def _symbol_intervals(eng: HSICEngine, tag: str, y_hat_nchw: torch.Tensor, ch: torch.Tensor, minmax: int) -> np.ndarray:
    """(n_pos * n_ch, 3) int32 coding intervals of every listed latent element, raster order, channel-minor."""
    lib = eng.lib
    M, K = eng.M, eng.K
    y_nhwc = y_hat_nchw.permute(0, 2, 3, 1).contiguous()
    n_pos = y_nhwc.shape[1] * y_nhwc.shape[2]
    out = torch.empty(n_pos * ch.numel(), 3, dtype=torch.int32, device=eng.dev)
    check(lib.masic_gmm_symbol_cdfs(eng.buf[f"{tag}.sigma"].data_ptr(), eng.buf[f"{tag}.mu"].data_ptr(),
                                    eng.buf[f"{tag}.wlogit"].data_ptr(), 1, M, K, n_pos, ch.data_ptr(), ch.numel(),
                                    minmax, SCALE_BOUND, y_nhwc.data_ptr(), None, out.data_ptr(), _stream()),
          "masic_gmm_symbol_cdfs")
    return out.cpu().numpy()


def compress(model, x1: torch.Tensor, x2: torch.Tensor, h_matrix: torch.Tensor, output_name, output_path: str = "",
             device=None) -> Dict:
    """HSIC.compress (MASIC.py:855-1158).  Batch 1, like the reference's file format."""
    if model.training:
        raise MasicError("HSIC.compress: eval mode only")
    if x1.shape[0] != 1:
        raise ValueError("HSIC.compress writes one stereo pair per file (MASIC.py:922: a single H, W header)")
    lib = _lib.load()
    _, _, H, W = x1.shape
    eng: HSICEngine = model.engine_for(1, H, W, x1.device)
    out = eng.run(x1, x2, h_matrix)
    M = model.M
    # ---- side information: z strings through the reference rANS (entropy_models.py:420-423)
    z_strings = []
    for tag, eb in (("L", model.entropy_bottleneck1), ("R", model.entropy_bottleneck2)):
        z = eng.buf[f"{tag}.z"].permute(0, 3, 1, 2).contiguous()
        z_strings.append(eb.compress(z))
    y_hats = {"L": out["y1_hat"], "R": out["y2_hat"]}
    flags, minmaxs = {}, {}
    for tag in ("L", "R"):
        flags[tag], minmaxs[tag] = _nonzero_channels(y_hats[tag])
    output1 = os.path.join(output_path, str(output_name) + ".npz")
    with open(output1, "wb") as f:
        f.write(np.array([H, W], dtype=np.uint16).tobytes())
        for tag, zs in zip(("L", "R"), z_strings):
            if len(zs[0]) > 65535:
                raise MasicError("z bitstream exceeds the uint16 length field of the reference header (MASIC.py:945)")
            f.write(np.array([len(zs[0]), minmaxs[tag]], dtype=np.uint16).tobytes())
            f.write(_pack_flags(flags[tag]))
            f.write(zs[0])
    # ---- y1, y2: all coding intervals in one kernel each, one range-coded stream
    start = time.time()
    ivs = []
    for tag in ("L", "R"):
        ch = torch.from_numpy(np.flatnonzero(flags[tag]).astype(np.int32)).to(eng.dev)
        if ch.numel():
            ivs.append(_symbol_intervals(eng, tag, y_hats[tag], ch, minmaxs[tag]))
    iv = np.ascontiguousarray(np.concatenate(ivs, 0)) if ivs else np.zeros((0, 3), np.int32)
    buf = np.empty(iv.shape[0] * 3 + 64, dtype=np.uint8)
    n_out = C.c_int64()
    check(lib.masic_range_encode(iv.ctypes.data, iv.shape[0], buf.ctypes.data, buf.size, C.byref(n_out)),
          "masic_range_encode")
    output2 = os.path.join(output_path, str(output_name) + ".bin")
    with open(output2, "wb") as f:
        f.write(buf[:n_out.value].tobytes())
    end = time.time()
    num_pixels = H * W * 2
    size1, size2 = os.path.getsize(output1), os.path.getsize(output2)
    ideal_bits = float(-(np.log2(iv[:, 1].astype(np.float64) / iv[:, 2])).sum()) if iv.shape[0] else 0.0
    return {
        "bpp_real": (size1 + size2) * 8 / num_pixels, "bpp_side": size1 * 8 / num_pixels, "enctime": end - start,
        "y1_hat": out["y1_hat"].clone(), "y2_hat": out["y2_hat"].clone(),
        "z1_hat": out["z1_hat"].clone(), "z2_hat": out["z2_hat"].clone(),
        "y_bits_ideal": ideal_bits, "y_bytes": size2, "n_symbols": int(iv.shape[0]),
    }


def _decode_view(eng: HSICEngine, tag: str, dec, flag: np.ndarray, minmax: int) -> torch.Tensor:
    """Sequential (raster) decode of one view's latents; returns y_hat NCHW fp32 and fills eng.buf[tag.y_rnd]."""
    lib = eng.lib
    M, K = eng.M, eng.K
    h16, w16 = eng.H // 16, eng.W // 16
    pm = _PixelModel(eng, tag)
    gmm_in = eng.buf[f"{tag}.gmm_in"]
    ypad = torch.zeros(1, h16 + 4, w16 + 4, M, dtype=torch.bfloat16, device=eng.dev)
    y_nhwc = torch.zeros(1, h16, w16, M, dtype=torch.float32, device=eng.dev)
    ch_np = np.flatnonzero(flag).astype(np.int32)
    n_ch = int(ch_np.size)
    if n_ch:
        ch = torch.from_numpy(ch_np).to(eng.dev)
        ch_long = ch.long()
        L1 = 2 * minmax + 2
        rows = torch.empty(n_ch, L1, dtype=torch.int32, device=eng.dev)
        rows_h = torch.empty(n_ch, L1, dtype=torch.int32).pin_memory()
        sym_h = np.empty(n_ch, dtype=np.int32)
        for h in range(h16):
            for w in range(w16):
                sig, mu, wl = pm.params_at(ypad, gmm_in, h, w)
                check(lib.masic_gmm_symbol_cdfs(sig.data_ptr(), mu.data_ptr(), wl.data_ptr(), 1, M, K, 1, ch.data_ptr(),
                                                n_ch, minmax, SCALE_BOUND, None, rows.data_ptr(), None, _stream()),
                      "masic_gmm_symbol_cdfs")
                rows_h.copy_(rows)                       # synchronises the pixel's kernels
                check(lib.masic_range_decode_rows(dec, rows_h.data_ptr(), n_ch, L1, sym_h.ctypes.data),
                      "masic_range_decode_rows")
                vals = torch.from_numpy(sym_h.astype(np.float32) - float(minmax)).to(eng.dev)
                y_nhwc[0, h, w, ch_long] = vals
                ypad[0, h + 2, w + 2, ch_long] = vals.to(torch.bfloat16)
    eng.buf[f"{tag}.y_rnd"].copy_(ypad[:, 2:-2, 2:-2, :])
    return y_nhwc.permute(0, 3, 1, 2).contiguous()


def decompress(model, x1: Optional[torch.Tensor], x2: Optional[torch.Tensor], h_matrix: torch.Tensor, output_name,
               output_path: str = "", device=None) -> Dict:
    """HSIC.decompress (MASIC.py:1161-1408).  x1 / x2 are only consulted for the device (the reference uses them
    for shapes, :1313); everything else comes from the two files and the homography."""
    if model.training:
        raise MasicError("HSIC.decompress: eval mode only")
    lib = _lib.load()
    M = model.M
    output1 = os.path.join(output_path, str(output_name) + ".npz")
    output2 = os.path.join(output_path, str(output_name) + ".bin")
    with open(output1, "rb") as f:
        H, W = (int(v) for v in np.frombuffer(f.read(4), dtype=np.uint16))
        hdr = []
        for _ in range(2):
            length, minmax = (int(v) for v in np.frombuffer(f.read(4), dtype=np.uint16))
            flag = np.unpackbits(np.frombuffer(f.read(M // 8), dtype=np.uint8))
            hdr.append((f.read(length), minmax, flag))
    dev = h_matrix.device if device is None else torch.device(device)
    if dev.type != "cuda":
        dev = x1.device if x1 is not None else torch.device("cuda:0")
    eng: HSICEngine = model.engine_for(1, H, W, dev)
    o = eng.out
    z_shape = (H // 64, W // 64)
    with torch.cuda.device(eng.dev):
        eng.Hm.copy_(h_matrix.reshape(1, 3, 3).to(eng.dev))
        # homography products and the mask-derived fusion weights (MASIC.py:1310-1312)
        eng.run_steps(lambda n: n.startswith("warp.prepare") or n.startswith("mask_") or n.startswith("mask2weights"))
        # side information -> hyper-synthesis parameters (:1190-1196)
        for tag, eb, (string, _, _) in zip(("L", "R"), (model.entropy_bottleneck1, model.entropy_bottleneck2), hdr):
            z_hat = eb.decompress([string], z_shape).to(eng.dev)
            o["z1_hat" if tag == "L" else "z2_hat"].copy_(z_hat)
            eng.buf[f"{tag}.zq"].copy_(z_hat.permute(0, 2, 3, 1))
            eng.run_steps(lambda n, t=tag: n.startswith(f"{t}.h_s."))
        with open(output2, "rb") as f:
            data = np.frombuffer(f.read(), dtype=np.uint8).copy()
        dec = C.c_void_p()
        check(lib.masic_range_decoder_create(data.ctypes.data, data.size, C.byref(dec)), "masic_range_decoder_create")
        try:
            start = time.time()
            o["y1_hat"].copy_(_decode_view(eng, "L", dec, hdr[0][2], hdr[0][1]))
            # left reconstruction, its warp, encoder1 on it and the mask-weighted prior term (:1303-1318)
            eng.run_steps(lambda n: n.startswith("L.g_s.") or n.startswith("L.x1_hat") or n == "R.warp(x1_hat)"
                          or n.startswith("R.g_a(enc1 on warped x1_hat)") or n.startswith("R.y1warp"))
            o["y2_hat"].copy_(_decode_view(eng, "R", dec, hdr[1][2], hdr[1][1]))
            eng.run_steps(lambda n: n.startswith("R.g_s.") or n.startswith("R.after_"))
            torch.cuda.synchronize(eng.dev)
            end = time.time()
        finally:
            lib.masic_range_decoder_destroy(dec)
    return {"x1_hat": o["x1_hat"].clone(), "x2_hat": o["x2_hat"].clone(), "y1_hat": o["y1_hat"].clone(),
            "y2_hat": o["y2_hat"].clone(), "z1_hat": o["z1_hat"].clone(), "z2_hat": o["z2_hat"].clone(),
            "dectime": end - start}
