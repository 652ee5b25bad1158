"""CPU restatement of the reference's entropy-model arithmetic (TEST INFRASTRUCTURE).

Plain numpy / torch-CPU fp32; every function cites the reference lines it restates
(paths relative to the reference root).  Pinned against fixtures generated from the
unmodified reference by tests/golden/make_golden.py (see tests/test_oracle_pinned.py).
"""
from __future__ import annotations

import ctypes
import math
from pathlib import Path
from typing import Sequence

import numpy as np
import scipy.stats
import torch
import torch.nn.functional as F

_ORACLE = Path(__file__).resolve().parent
LIKELIHOOD_BOUND = 1e-9          # entropy_models.py:66
SCALE_BOUND = 0.11               # entropy_models.py:445,728
TAIL_MASS = 1e-9                 # entropy_models.py:256,446


# --------------------------------------------------------------------------- pmf -> cdf
def pmf_to_quantized_cdf(pmf: Sequence[float], precision: int = 16) -> np.ndarray:
    """compressai/cpp_exts/ops/ops.cpp:40-109 — float pmf -> strictly increasing uint32 cdf.

    round(p * 2^precision) in float32, renormalise by integer division so the total is
    <= 2^precision, prefix-sum, force the last entry to 2^precision, then for every empty
    bin steal one count from the smallest bin that has more than one (first such bin wins).
    """
    pmf32 = np.asarray(pmf, dtype=np.float32)
    if pmf32.size and (np.any(pmf32 < 0) or not np.all(np.isfinite(pmf32))):
        raise ValueError("Invalid `pmf`, non-finite or negative element found")
    scale = np.float32(1 << precision)
    # std::round on float = round half away from zero
    # the float32 product is exact (power-of-two scale); widen to float64 so that +0.5 is exact too
    rounded = np.floor((pmf32 * scale).astype(np.float64) + 0.5).astype(np.int64)
    freq = [0] + [int(v) for v in rounded]
    total = sum(freq)
    if total == 0:
        raise ValueError("Invalid `pmf`: at least one element must have a non-zero probability.")
    freq = [((1 << precision) * f) // total for f in freq]
    cdf = list(np.cumsum(np.asarray(freq, dtype=np.int64)))
    cdf = [int(c) for c in cdf]
    cdf[-1] = 1 << precision
    n = len(cdf)
    for i in range(n - 1):
        if cdf[i] == cdf[i + 1]:
            best_freq, best = 1 << 40, -1
            for j in range(n - 1):
                f = cdf[j + 1] - cdf[j]
                if 1 < f < best_freq:
                    best_freq, best = f, j
            assert best != -1
            if best < i:
                for j in range(best + 1, i + 1):
                    cdf[j] -= 1
            else:
                for j in range(i + 1, best + 1):
                    cdf[j] += 1
    return np.asarray(cdf, dtype=np.uint32)


_clib = None


def pmf_to_quantized_cdf_c(pmf: Sequence[float], precision: int = 16) -> np.ndarray:
    """Same algorithm through oracle/pmf_to_cdf.c (fast path for big tables)."""
    global _clib
    if _clib is None:
        so = _ORACLE / "_build" / "libmasic_oracle.so"
        if not so.exists():
            raise FileNotFoundError(f"{so}: run `make -C oracle port`")
        _clib = ctypes.CDLL(str(so))
        _clib.masic_oracle_pmf_to_quantized_cdf.restype = ctypes.c_int
        _clib.masic_oracle_pmf_to_quantized_cdf.argtypes = [
            ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)]
    p = np.ascontiguousarray(pmf, dtype=np.float32)
    out = np.zeros(p.size + 1, dtype=np.uint32)
    rc = _clib.masic_oracle_pmf_to_quantized_cdf(
        p.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), p.size, precision,
        out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    if rc != 0:
        raise ValueError(f"pmf_to_quantized_cdf: invalid pmf (code {rc})")
    return out


def pmf_rows_to_cdf(pmf: torch.Tensor, tail_mass: torch.Tensor, pmf_length: torch.Tensor,
                    max_length: int, precision: int = 16, use_c: bool = True) -> torch.Tensor:
    """entropy_models.py:136-142 — one cdf row per table entry: pmf[:len] ++ tail_mass."""
    fn = pmf_to_quantized_cdf_c if use_c else pmf_to_quantized_cdf
    cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
    for i in range(len(pmf_length)):
        prob = torch.cat((pmf[i, : int(pmf_length[i])], tail_mass[i].reshape(-1)), dim=0)
        row = fn(prob.numpy(), precision)
        cdf[i, : row.size] = torch.from_numpy(row.astype(np.int64)).to(torch.int32)
    return cdf


# --------------------------------------------------------------------------- quantisation
def quantize_symbols(x: torch.Tensor, means: torch.Tensor | None = None) -> torch.Tensor:
    """entropy_models.py:112-125 ('symbols'): int32(round_half_even(x - means))."""
    v = x.clone()
    if means is not None:
        v = v - means
    return torch.round(v).to(torch.int32)


def quantize_dequantize(x: torch.Tensor, means: torch.Tensor | None = None) -> torch.Tensor:
    """entropy_models.py:112-121 ('dequantize'): round(x - means) + means."""
    v = x.clone()
    if means is not None:
        v = v - means
    v = torch.round(v)
    if means is not None:
        v = v + means
    return v


# --------------------------------------------------------------------------- EntropyBottleneck
class EBParams:
    """The learnable state of one EntropyBottleneck (entropy_models.py:267-296)."""

    def __init__(self, matrices, biases, factors, quantiles):
        self.matrices = [m.detach().float() for m in matrices]
        self.biases = [b.detach().float() for b in biases]
        self.factors = [f.detach().float() for f in factors]
        self.quantiles = quantiles.detach().float()

    @classmethod
    def from_state_dict(cls, sd, prefix: str):
        n = 0
        while f"{prefix}_matrices.{n}" in sd:
            n += 1
        return cls([sd[f"{prefix}_matrices.{i}"] for i in range(n)],
                   [sd[f"{prefix}_biases.{i}"] for i in range(n)],
                   [sd[f"{prefix}_factors.{i}"] for i in range(n - 1)],
                   sd[f"{prefix}quantiles"])

    @property
    def channels(self):
        return self.quantiles.shape[0]

    def medians(self):
        return self.quantiles[:, :, 1:2]          # (C,1,1)  entropy_models.py:298-300


def eb_logits_cumulative(p: EBParams, v: torch.Tensor) -> torch.Tensor:
    """entropy_models.py:350-369: v (C,1,L) -> logits (C,1,L)."""
    logits = v
    for i, m in enumerate(p.matrices):
        logits = torch.matmul(F.softplus(m), logits)
        logits = logits + p.biases[i]
        if i < len(p.factors):
            logits = logits + torch.tanh(p.factors[i]) * torch.tanh(logits)
    return logits


def eb_likelihood(p: EBParams, v: torch.Tensor) -> torch.Tensor:
    """entropy_models.py:372-382."""
    lower = eb_logits_cumulative(p, v - 0.5)
    upper = eb_logits_cumulative(p, v + 0.5)
    sign = -torch.sign(lower + upper)
    return torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))


def eb_forward(p: EBParams, z: torch.Tensor):
    """entropy_models.py:384-411 (eval): returns (z_hat, likelihood), both (N,C,H,W)."""
    zc = z.permute(1, 2, 3, 0).contiguous()
    shape = zc.shape
    vals = zc.reshape(shape[0], 1, -1)
    q = quantize_dequantize(vals, p.medians())
    lik = torch.clamp_min(eb_likelihood(p, q), LIKELIHOOD_BOUND)
    back = lambda t: t.reshape(shape).permute(3, 0, 1, 2).contiguous()  # noqa: E731
    return back(q), back(lik)


def eb_symbols(p: EBParams, z: torch.Tensor) -> torch.Tensor:
    """entropy_models.py:420-423 + :174: int32(round(z - median_c))."""
    return quantize_symbols(z, p.medians().view(1, -1, 1, 1))


def eb_indexes(size) -> torch.Tensor:
    """entropy_models.py:413-418: index = channel id."""
    n, c, h, w = size
    return torch.arange(c, dtype=torch.int32).view(1, -1, 1, 1).repeat(n, 1, h, w)


def eb_tables(p: EBParams, precision: int = 16, use_c: bool = True):
    """entropy_models.py:302-343 -> (_offset, _quantized_cdf, _cdf_length), all int32."""
    med = p.quantiles[:, 0, 1]
    minima = torch.clamp(torch.ceil(med - p.quantiles[:, 0, 0]).int(), min=0)
    maxima = torch.clamp(torch.ceil(p.quantiles[:, 0, 2] - med).int(), min=0)
    offset = -minima
    pmf_start = med - minima
    pmf_length = maxima + minima + 1
    max_length = int(pmf_length.max())
    samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
    lower = eb_logits_cumulative(p, samples - 0.5)
    upper = eb_logits_cumulative(p, samples + 0.5)
    sign = -torch.sign(lower + upper)
    pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
    tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
    cdf = pmf_rows_to_cdf(pmf, tail, pmf_length, max_length, precision, use_c)
    return offset.to(torch.int32), cdf, (pmf_length + 2).to(torch.int32)


def eb_aux_loss(p: EBParams, target: torch.Tensor) -> torch.Tensor:
    """entropy_models.py:345-348."""
    return torch.abs(eb_logits_cumulative(p, p.quantiles) - target).sum()


# --------------------------------------------------------------------------- Gaussian conditional
def std_cumulative(x: torch.Tensor) -> torch.Tensor:
    """entropy_models.py:484-489: Phi(x) = 0.5 * erfc(-x / sqrt(2))."""
    return 0.5 * torch.erfc(float(-(2 ** -0.5)) * x)


def default_scale_table(lo: float = 0.11, hi: float = 256.0, levels: int = 64):
    """compressai/models/google.py:195-201 (get_scale_table)."""
    return [float(v) for v in torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))]


def gc_likelihood(values: torch.Tensor, scales: torch.Tensor, means: torch.Tensor | None = None,
                  bound: float = SCALE_BOUND) -> torch.Tensor:
    """entropy_models.py:528-544 (no likelihood floor)."""
    v = values if means is None else values - means
    s = torch.clamp_min(scales, bound)
    v = torch.abs(v)
    return std_cumulative((0.5 - v) / s) - std_cumulative((-0.5 - v) / s)


def gc_forward(y, scales, means=None, bound: float = SCALE_BOUND):
    """entropy_models.py:546-554 (eval)."""
    q = quantize_dequantize(y, means)
    return q, torch.clamp_min(gc_likelihood(q, scales, means, bound), LIKELIHOOD_BOUND)


def gc_build_indexes(scales: torch.Tensor, table: Sequence[float], bound: float = SCALE_BOUND):
    """entropy_models.py:556-562: (L-1) - #{s in table[:-1] : max(scale, bound) <= s}."""
    s = torch.clamp_min(scales, bound)
    t = torch.tensor(list(table), dtype=torch.float32)
    idx = torch.full(s.shape, len(t) - 1, dtype=torch.int32)
    for v in t[:-1]:
        idx -= (s <= v).int()
    return idx


def gc_tables(table: Sequence[float], tail_mass: float = TAIL_MASS, precision: int = 16,
              use_c: bool = True):
    """entropy_models.py:504-526 -> (_offset, _quantized_cdf, _cdf_length)."""
    st = torch.tensor([float(s) for s in table], dtype=torch.float32)
    mult = -scipy.stats.norm.ppf(tail_mass / 2)
    center = torch.ceil(st * mult).int()
    length = 2 * center + 1
    max_length = int(length.max())
    samples = torch.abs(torch.arange(max_length).int() - center[:, None]).float()
    sc = st.unsqueeze(1)
    upper = std_cumulative((0.5 - samples) / sc)
    lower = std_cumulative((-0.5 - samples) / sc)
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    cdf = pmf_rows_to_cdf(pmf, tail, length, max_length, precision, use_c)
    return (-center).to(torch.int32), cdf, (length + 2).to(torch.int32)


# --------------------------------------------------------------------------- K-component mixture
def gmm_likelihood(y_hat: torch.Tensor, sigma: torch.Tensor, mu: torch.Tensor, w: torch.Tensor,
                   K: int, bound: float = SCALE_BOUND) -> torch.Tensor:
    """entropy_models.py:808-846: sum_k w_k [Phi((.5-|v-mu_k|)/s_k) - Phi((-.5-|v-mu_k|)/s_k)],
    parameters k-major along channels (ch = k*M + m)."""
    M = y_hat.shape[1]
    lik = None
    for k in range(K):
        sl = slice(M * k, M * (k + 1))
        term = gc_likelihood(y_hat, sigma[:, sl], mu[:, sl], bound) * w[:, sl]
        lik = term if lik is None else lik + term
    return lik


def gmm_forward(y, sigma, mu, w, K: int, bound: float = SCALE_BOUND):
    """entropy_models.py:849-858 (eval): y_hat = round(y) (means=None), lik floored at 1e-9."""
    y_hat = torch.round(y)
    return y_hat, torch.clamp_min(gmm_likelihood(y_hat, sigma, mu, w, K, bound), LIKELIHOOD_BOUND)


def gmm_symbol_cdf(sigma_k, mu_k, w_k, minmax: int, bound: float = SCALE_BOUND) -> np.ndarray:
    """coremasic/mywork/MASIC.py:999-1043: the per-symbol integer CDF HSIC.compress hands to
    the range coder.  Support [0, 2*minmax], means shifted by +minmax, pmf clipped to
    [1/65536, 1], renormalised to 65536 with np.round, float32 prefix sum."""
    samples = torch.arange(0, 2 * minmax + 1, dtype=torch.float32)
    pmf = None
    for k in range(len(sigma_k)):
        v = torch.abs(samples - (mu_k[k] + minmax))
        s = torch.clamp_min(sigma_k[k], bound)
        t = (std_cumulative((0.5 - v) / s) - std_cumulative((-0.5 - v) / s)) * w_k[k]
        pmf = t if pmf is None else pmf + t
    p = pmf.numpy().astype(np.float32)
    clip = np.clip(p, 1.0 / 65536, 1.0)
    clip = np.round(clip / np.sum(clip) * 65536)
    cdf = np.add.accumulate(clip)
    return np.asarray([0] + [int(c) for c in cdf], dtype=np.int64)
