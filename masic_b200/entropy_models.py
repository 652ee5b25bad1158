"""compressai.entropy_models surface (EntropyBottleneck, GaussianConditional,
GaussianMixtureConditional[_gf]) on top of the sm_100a kernels.

Constructor signatures, parameter/buffer names, error messages and call semantics follow
compressai/entropy_models/entropy_models.py of the reference (line numbers cited per
method).  Quantise / likelihood / CDF-index computation run in CUDA (masic_b200/csrc/
entropy.cu); the integer CDF construction runs in the library's host code (csrc/cdf.cu);
rANS serialisation is host code outside the GPU hot path (BASELINE.json's north-star): the
library's native coder (csrc/rans.cpp) writes the byte format of the reference's `compressai.ans`
extension and takes int32 buffers instead of Python lists.
"""
from __future__ import annotations

import importlib
import os
from typing import List, Optional

import numpy as np
import scipy.stats
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import ops
from . import torch_ops  # noqa: F401  (registers torch.ops.masic_b200.*)
from ._lib import MasicError
from .layers import LowerBound

__all__ = ["EntropyModel", "EntropyBottleneck", "GaussianConditional", "GaussianMixtureConditional",
           "GaussianMixtureConditional_gf", "pmf_to_quantized_cdf"]


def pmf_to_quantized_cdf(pmf: Tensor, precision: int = 16) -> Tensor:
    """entropy_models.py:50-53."""
    return torch.IntTensor(ops.pmf_to_quantized_cdf(pmf.tolist(), precision))


def _load_ans():
    """The rANS coder behind EntropyModel.compress / decompress.  Default: the library's own native coder
    (masic_b200/rans.py over `masic_rans_*`, csrc/rans.cpp), which writes the byte format of the reference's
    `compressai.ans` extension and takes int32 buffers instead of Python lists.  $MASIC_ANS_MODULE names another
    module with the `compressai.ans` surface (e.g. the reference's own extension where it is installed); nothing
    here looks under oracle/."""
    name = os.environ.get("MASIC_ANS_MODULE")
    if name:
        return importlib.import_module(name)
    from . import rans
    return rans


class _EntropyCoder:
    """entropy_models.py:13-42 — proxy to the rANS coder."""

    def __init__(self, method: str):
        if not isinstance(method, str):
            raise ValueError(f'Invalid method type "{type(method)}"')
        if method != "ans":
            raise ValueError(f'Unknown entropy coder "{method}" (available: ans)')
        self._encoder = None
        self._decoder = None
        self.native = False

    def _ensure(self):
        if self._encoder is None:
            ans = _load_ans()
            self._encoder, self._decoder = ans.RansEncoder(), ans.RansDecoder()
            self.native = hasattr(self._decoder, "decode_with_indexes_array")

    def encode_with_indexes(self, *a, **k):
        self._ensure()
        return self._encoder.encode_with_indexes(*a, **k)

    def decode_with_indexes(self, *a, **k):
        self._ensure()
        return self._decoder.decode_with_indexes(*a, **k)


class EntropyModel(nn.Module):
    """entropy_models.py:56-239."""

    def __init__(self, likelihood_bound: float = 1e-9, entropy_coder: Optional[str] = None,
                 entropy_coder_precision: int = 16):
        super().__init__()
        self.entropy_coder = _EntropyCoder("ans" if entropy_coder is None else entropy_coder)
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        self.likelihood_bound = float(likelihood_bound)
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    def forward(self, *args):
        raise NotImplementedError()

    def _quantize(self, inputs: Tensor, mode: str, means: Optional[Tensor] = None) -> Tensor:
        """entropy_models.py:98-125.  'noise' (training) is outside the inference path."""
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            raise MasicError("quantisation mode 'noise' is training-only; masic_b200 implements inference")
        if means is not None:
            means = means.expand_as(inputs)
        return torch.ops.masic_b200.quantize(inputs, means, mode == "symbols")

    @staticmethod
    def _dequantize(inputs: Tensor, means: Optional[Tensor] = None) -> Tensor:
        """entropy_models.py:127-134."""
        if means is not None:
            outputs = inputs.type_as(means)
            outputs = outputs + means
        else:
            outputs = inputs.float()
        return outputs

    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        """entropy_models.py:136-142 (whole table in one native call)."""
        return ops.pmf_table_to_cdf(pmf, tail_mass, pmf_length, int(max_length), self.entropy_coder_precision)

    def _check_cdf_size(self):
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if len(self._quantized_cdf.size()) != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")

    def _check_offsets_size(self):
        if self._offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if len(self._offset.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._offset.size()}")

    def _check_cdf_length(self):
        if self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")
        if len(self._cdf_length.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._cdf_length.size()}")

    # The coder tables only change in update() / update_scale_table() / load_state_dict(): their host-side copies
    # (an int32 TableSet for the native coder, Python lists for a `compressai.ans`-style module — the reference
    # converts them on every call, entropy_models.py:192-194) are cached and dropped EXPLICITLY whenever one of the
    # three buffers is assigned or loaded; staleness is never inferred from versions or pointers.
    _TABLE_BUFFERS = ("_quantized_cdf", "_cdf_length", "_offset")

    def __setattr__(self, name, value):
        if name in EntropyModel._TABLE_BUFFERS:
            self.__dict__.pop("_tbl_cache", None)
        super().__setattr__(name, value)

    def _load_from_state_dict(self, *args, **kwargs):
        self.__dict__.pop("_tbl_cache", None)
        return super()._load_from_state_dict(*args, **kwargs)

    def invalidate_tables(self):
        """Call after editing `_quantized_cdf` / `_cdf_length` / `_offset` IN PLACE."""
        self.__dict__.pop("_tbl_cache", None)

    def _coder_tables(self):
        self.entropy_coder._ensure()
        cached = self.__dict__.get("_tbl_cache")
        if cached is None:
            if self.entropy_coder.native:
                from .rans import TableSet
                cached = (TableSet(self._quantized_cdf, self._cdf_length.reshape(-1), self._offset.reshape(-1)),)
            else:
                cached = (self._quantized_cdf.tolist(), self._cdf_length.reshape(-1).int().tolist(),
                          self._offset.reshape(-1).int().tolist())
            self.__dict__["_tbl_cache"] = cached
        return cached

    def compress(self, inputs: Tensor, indexes: Tensor, means: Optional[Tensor] = None) -> List[bytes]:
        """entropy_models.py:165-196: symbols on the GPU; serialisation by the rANS coder on int32 buffers."""
        symbols = self._quantize(inputs, "symbols", means)
        if len(inputs.size()) != 4:
            raise ValueError("Invalid `inputs` size. Expected a 4-D tensor.")
        if inputs.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        tables = self._coder_tables()
        sym_cpu = symbols.to(torch.int32).cpu()
        idx_cpu = indexes.to(torch.int32).cpu()
        strings = []
        for i in range(sym_cpu.size(0)):
            if self.entropy_coder.native:
                strings.append(self.entropy_coder.encode_with_indexes(
                    sym_cpu[i].reshape(-1).numpy(), idx_cpu[i].reshape(-1).numpy(), *tables))
            else:
                strings.append(self.entropy_coder.encode_with_indexes(
                    sym_cpu[i].reshape(-1).tolist(), idx_cpu[i].reshape(-1).tolist(), *tables))
        return strings

    def decompress(self, strings, indexes: Tensor, means: Optional[Tensor] = None) -> Tensor:
        """entropy_models.py:199-239."""
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if not len(strings) == indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if len(indexes.size()) != 4:
            raise ValueError("Invalid `indexes` size. Expected a 4-D tensor.")
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        if means is not None:
            if means.size()[:-2] != indexes.size()[:-2]:
                raise ValueError("Invalid means or indexes parameters")
            if means.size() != indexes.size() and (means.size(2) != 1 or means.size(3) != 1):
                raise ValueError("Invalid means parameters")
        tables = self._coder_tables()
        idx_cpu = indexes.to(torch.int32).cpu()
        outputs = torch.empty(indexes.size(), dtype=torch.int32)
        for i, s in enumerate(strings):
            if self.entropy_coder.native:
                values = torch.from_numpy(self.entropy_coder._decoder.decode_with_indexes_array(
                    s, idx_cpu[i].reshape(-1).numpy(), *tables))
            else:
                values = torch.tensor(self.entropy_coder.decode_with_indexes(
                    s, idx_cpu[i].reshape(-1).tolist(), *tables), dtype=torch.int32)
            outputs[i] = values.reshape(outputs[i].size())
        outputs = outputs.to(self._quantized_cdf.device if means is None else means.device)
        return self._dequantize(outputs, means)


class EntropyBottleneck(EntropyModel):
    """entropy_models.py:242-430."""

    def __init__(self, channels: int, *args, tail_mass: float = 1e-9, init_scale: float = 10,
                 filters=(3, 3, 3, 3), **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        self._biases = nn.ParameterList()
        self._factors = nn.ParameterList()
        self._matrices = nn.ParameterList()
        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        channels = self.channels
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filters[i + 1]))
            matrix = torch.Tensor(channels, filters[i + 1], filters[i])
            matrix.data.fill_(init)
            self._matrices.append(nn.Parameter(matrix))
            bias = torch.Tensor(channels, filters[i + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self._biases.append(nn.Parameter(bias))
            if i < len(self.filters):
                factor = torch.Tensor(channels, filters[i + 1], 1)
                nn.init.zeros_(factor)
                self._factors.append(nn.Parameter(factor))
        self.quantiles = nn.Parameter(torch.Tensor(channels, 1, 3))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles.data = init.repeat(self.quantiles.size(0), 1, 1)
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    def _medians(self) -> Tensor:
        return self.quantiles[:, :, 1:2]

    def _logits_cumulative(self, inputs: Tensor, stop_gradient: bool) -> Tensor:
        """entropy_models.py:350-369 — used by update()/loss() (cold, table-sized inputs); kept on the
        same torch ops as the reference so the CDF tables (hence the bitstream) are bit-identical."""
        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix = self._matrices[i]
            if stop_gradient:
                matrix = matrix.detach()
            logits = torch.matmul(F.softplus(matrix), logits)
            bias = self._biases[i]
            if stop_gradient:
                bias = bias.detach()
            logits = logits + bias
            if i < len(self._factors):
                factor = self._factors[i]
                if stop_gradient:
                    factor = factor.detach()
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def update(self, force: bool = False) -> None:
        """entropy_models.py:302-343.  The pmf is evaluated on the CPU with the reference's torch ops
        (1-ulp differences in sigmoid/softplus would flip table entries and with them the
        bitstream — SURVEY §7); the integer part runs in the library's native host code."""
        if self._offset.numel() > 0 and not force:
            return
        dev = self.quantiles.device
        with torch.no_grad():
            cpu = EntropyBottleneck.__new__(EntropyBottleneck)
            nn.Module.__init__(cpu)
            cpu.filters = self.filters
            cpu._matrices = [m.detach().cpu() for m in self._matrices]
            cpu._biases = [b.detach().cpu() for b in self._biases]
            cpu._factors = [f.detach().cpu() for f in self._factors]
            q = self.quantiles.detach().cpu()
            medians = q[:, 0, 1]
            minima = torch.clamp(torch.ceil(medians - q[:, 0, 0]).int(), min=0)
            maxima = torch.clamp(torch.ceil(q[:, 0, 2] - medians).int(), min=0)
            offset = -minima
            pmf_start = medians - minima
            pmf_length = maxima + minima + 1
            max_length = int(pmf_length.max())
            samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
            half = float(0.5)
            lower = EntropyBottleneck._logits_cumulative(cpu, samples - half, stop_gradient=True)
            upper = EntropyBottleneck._logits_cumulative(cpu, samples + half, stop_gradient=True)
            sign = -torch.sign(lower + upper)
            pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
            tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
            cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        self._offset = offset.to(torch.int32).to(dev)
        self._quantized_cdf = cdf.to(dev)
        self._cdf_length = (pmf_length + 2).to(torch.int32).to(dev)

    def loss(self) -> Tensor:
        """entropy_models.py:345-348."""
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    def forward(self, x: Tensor):
        """entropy_models.py:384-411 (eval): one fused kernel instead of ~60 ATen launches."""
        if self.training:
            raise MasicError("EntropyBottleneck: training mode (additive noise) is not on the sm_100a path")
        return torch.ops.masic_b200.eb_forward(x, list(self._matrices), list(self._biases), list(self._factors),
                                               self.quantiles)

    @staticmethod
    def _build_indexes(size):
        """entropy_models.py:413-418."""
        n, c, h, w = size
        return torch.arange(c).view(1, -1, 1, 1).int().repeat(n, 1, h, w)

    def compress(self, x: Tensor):
        """entropy_models.py:420-423."""
        indexes = self._build_indexes(x.size())
        medians = self._medians().detach().view(1, -1, 1, 1)
        return super().compress(x, indexes, medians)

    def decompress(self, strings, size):
        """entropy_models.py:425-430."""
        output_size = (len(strings), self._quantized_cdf.size(0), size[0], size[1])
        indexes = self._build_indexes(output_size)
        medians = self._medians().detach().view(1, -1, 1, 1)
        return super().decompress(strings, indexes, medians)


class _GaussianBase(EntropyModel):
    """Shared state of the Gaussian family (entropy_models.py:442-526 / 575-657 / 722-804)."""

    def _init_gaussian(self, scale_table, scale_bound, tail_mass):
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]) if scale_bound is not None else None)
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            self.lower_bound_scale = LowerBound(self.scale_table[0])
        elif scale_bound > 0:
            self.lower_bound_scale = LowerBound(scale_bound)
        else:
            raise ValueError("Invalid parameters")

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    def _standardized_cumulative(self, inputs: Tensor) -> Tensor:
        half = float(0.5)
        const = float(-(2 ** -0.5))
        return half * torch.erfc(const * inputs)

    @staticmethod
    def _standardized_quantile(quantile):
        return scipy.stats.norm.ppf(quantile)

    def _bound(self) -> float:
        return float(self.lower_bound_scale.bound)

    def update_scale_table(self, scale_table, force: bool = False):
        if self._offset.numel() > 0 and not force:
            return
        self.scale_table = self._prepare_scale_table(scale_table).to(self.scale_table.device)
        self.update()

    def update(self):
        """entropy_models.py:504-526 — pmf on the CPU with the reference's torch ops, integer part native."""
        dev = self.scale_table.device
        st = self.scale_table.detach().cpu()
        multiplier = -self._standardized_quantile(self.tail_mass / 2)
        pmf_center = torch.ceil(st * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = torch.max(pmf_length).item()
        samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
        samples_scale = st.unsqueeze(1).float()
        upper = self._standardized_cumulative((.5 - samples) / samples_scale)
        lower = self._standardized_cumulative((-.5 - samples) / samples_scale)
        pmf = upper - lower
        tail_mass = 2 * lower[:, :1]
        cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        self._quantized_cdf = cdf.to(dev)
        self._offset = (-pmf_center).to(torch.int32).to(dev)
        self._cdf_length = (pmf_length + 2).to(torch.int32).to(dev)

    def build_indexes(self, scales: Tensor) -> Tensor:
        """entropy_models.py:556-562 — one kernel (binary search) instead of 63 compare+sub launches."""
        return torch.ops.masic_b200.gc_build_indexes(scales, self.scale_table, self._bound())


class GaussianConditional(_GaussianBase):
    """entropy_models.py:433-562."""

    def __init__(self, scale_table, *args, scale_bound: float = 0.11, tail_mass: float = 1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        self._init_gaussian(scale_table, scale_bound, tail_mass)

    def _likelihood(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None) -> Tensor:
        """entropy_models.py:528-544 on already-quantised inputs (no floor)."""
        v = inputs if means is None else inputs - means
        s = torch.max(scales, self.lower_bound_scale.bound)
        v = torch.abs(v)
        return self._standardized_cumulative((0.5 - v) / s) - self._standardized_cumulative((-0.5 - v) / s)

    def forward(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None):
        """entropy_models.py:546-554 (eval): quantise + likelihood (+floor) fused."""
        if self.training:
            raise MasicError("GaussianConditional: training mode is not on the sm_100a path")
        y_hat, lik = torch.ops.masic_b200.gc_likelihood(inputs, scales, None if means is None else means.expand_as(inputs),
                                                        self._bound())
        if not self.use_likelihood_bound:
            lik = self._likelihood(y_hat, scales, means)
        return y_hat, lik


class GaussianMixtureConditional_gf(_GaussianBase):
    """entropy_models.py:713-858 — the K-component, per-pixel-weight mixture HSIC instantiates
    (MASIC.py:658-659)."""

    def __init__(self, K, scale_table=None, mean_table=None, weight_table=None, *args,
                 scale_bound: float = 0.11, tail_mass: float = 1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        self.K = K
        self._init_gaussian(scale_table, scale_bound, tail_mass)

    def forward(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None,
                weights: Optional[Tensor] = None):
        """entropy_models.py:849-858 (eval): y_hat = round(y) (means=None there), mixture likelihood."""
        if self.training:
            raise MasicError("GaussianMixtureConditional_gf: training mode is not on the sm_100a path")
        if self.K != 5:
            raise MasicError("the fused mixture kernel is specialised for K = 5 (MASIC.py:653)")
        return torch.ops.masic_b200.gmm_likelihood(inputs, scales, means, weights, self._bound())


class GaussianMixtureConditional(GaussianMixtureConditional_gf):
    """entropy_models.py:566-710 — same arithmetic (weights broadcast per channel by the caller)."""
