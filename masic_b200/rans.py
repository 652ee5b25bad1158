"""`compressai.ans` surface (RansEncoder / BufferedRansEncoder / RansDecoder) on the library's native rANS coder.

The reference reaches its pybind11 module with Python lists (compressai/entropy_models/entropy_models.py:186-195:
`.tolist()` of the symbols, the indexes and the whole `_quantized_cdf` table on every call — SURVEY a12 measured
0.32 s per call for that conversion alone).  The classes below accept those same lists, so code written against
`compressai.ans` runs unchanged, and also int32 numpy arrays / CPU tensors, which go to the coder without any
conversion (`masic_rans_*` in include/masic_b200.h take plain int32 buffers).  Byte strings are identical to the
reference extension's (tests/test_rans_cpu.py).  Host code: this is serialisation, outside the GPU hot path.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import check

__all__ = ["BufferedRansEncoder", "RansEncoder", "RansDecoder", "TableSet"]


def _i32(a) -> np.ndarray:
    if hasattr(a, "detach"):                       # torch tensor (CPU)
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.int32).reshape(-1)


class TableSet:
    """(cdfs, cdf_sizes, offsets) as contiguous int32 buffers.  Built once per update() by the entropy models;
    built on the fly from lists (ragged rows are zero-padded) for the `compressai.ans` list API."""

    __slots__ = ("cdfs", "sizes", "offsets", "n_tables", "pitch")

    def __init__(self, cdfs, cdf_sizes, offsets):
        if hasattr(cdfs, "detach"):
            cdfs = cdfs.detach().cpu().numpy()
        if isinstance(cdfs, np.ndarray):
            tab = np.ascontiguousarray(cdfs, dtype=np.int32)
            if tab.ndim != 2:
                raise ValueError(f"Invalid CDF size {tab.shape}")
        else:
            rows = [np.asarray(r, dtype=np.int32).reshape(-1) for r in cdfs]
            pitch = max((r.size for r in rows), default=0)
            tab = np.zeros((len(rows), pitch), dtype=np.int32)
            for i, r in enumerate(rows):
                tab[i, :r.size] = r
        self.cdfs, self.sizes, self.offsets = tab, _i32(cdf_sizes), _i32(offsets)
        self.n_tables, self.pitch = tab.shape
        if self.sizes.size != self.n_tables or self.offsets.size != self.n_tables:
            raise ValueError("cdfs, cdfs_sizes and offsets must describe the same number of tables")

    def args(self) -> Tuple:
        return self.cdfs.ctypes.data, self.n_tables, self.pitch, self.sizes.ctypes.data, self.offsets.ctypes.data


def _tables(cdfs, cdf_sizes, offsets) -> TableSet:
    return cdfs if isinstance(cdfs, TableSet) else TableSet(cdfs, cdf_sizes, offsets)


class BufferedRansEncoder:
    """rans_interface.hpp:45-64."""

    def __init__(self):
        self._lib = _lib.load()
        h = C.c_void_p()
        check(self._lib.masic_rans_encoder_create(C.byref(h)), "masic_rans_encoder_create")
        self._h = h

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes=None, offsets=None) -> None:
        sym, idx = _i32(symbols), _i32(indexes)
        if sym.size != idx.size:
            raise ValueError("`symbols` and `indexes` should have the same size.")
        t = _tables(cdfs, cdfs_sizes, offsets)
        check(self._lib.masic_rans_encoder_push(self._h, sym.ctypes.data, idx.ctypes.data, sym.size, *t.args()),
              "masic_rans_encoder_push")

    def flush(self) -> bytes:
        p, n = C.c_void_p(), C.c_int64()
        check(self._lib.masic_rans_encoder_flush(self._h, C.byref(p), C.byref(n)), "masic_rans_encoder_flush")
        return C.string_at(p.value, n.value)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.masic_rans_encoder_destroy(h)
            self._h = None


class RansEncoder:
    """rans_interface.hpp:66-81."""

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes=None, offsets=None) -> bytes:
        enc = BufferedRansEncoder()
        enc.encode_with_indexes(symbols, indexes, cdfs, cdfs_sizes, offsets)
        return enc.flush()


class RansDecoder:
    """rans_interface.hpp:83-114.  The list API returns lists like the reference; `decode_array` returns the int32
    numpy array the coder filled (no list round trip)."""

    def __init__(self):
        self._lib = _lib.load()
        self._h = None

    def set_stream(self, stream: bytes) -> None:
        self._close()
        h = C.c_void_p()
        buf = bytes(stream)
        check(self._lib.masic_rans_decoder_create(buf, len(buf), C.byref(h)), "masic_rans_decoder_create")
        self._h = h

    def decode_array(self, indexes, cdfs, cdfs_sizes=None, offsets=None) -> np.ndarray:
        if self._h is None:
            raise ValueError("RansDecoder: set_stream() first")
        idx = _i32(indexes)
        t = _tables(cdfs, cdfs_sizes, offsets)
        out = np.empty(idx.size, dtype=np.int32)
        check(self._lib.masic_rans_decoder_decode(self._h, idx.ctypes.data, idx.size, *t.args(), out.ctypes.data),
              "masic_rans_decoder_decode")
        return out

    def decode_stream(self, indexes, cdfs, cdfs_sizes=None, offsets=None):
        return self.decode_array(indexes, cdfs, cdfs_sizes, offsets).tolist()

    def decode_with_indexes(self, encoded: bytes, indexes, cdfs, cdfs_sizes=None, offsets=None):
        self.set_stream(encoded)
        try:
            return self.decode_stream(indexes, cdfs, cdfs_sizes, offsets)
        finally:
            self._close()

    def decode_with_indexes_array(self, encoded: bytes, indexes, cdfs, cdfs_sizes=None, offsets=None) -> np.ndarray:
        self.set_stream(encoded)
        try:
            return self.decode_array(indexes, cdfs, cdfs_sizes, offsets)
        finally:
            self._close()

    def _close(self):
        if getattr(self, "_h", None):
            self._lib.masic_rans_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        self._close()
