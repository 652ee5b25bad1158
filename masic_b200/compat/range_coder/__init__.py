"""`range_coder` (PyPI; imported at MASIC.py:16, used at :958,1048,1125,1221,1298,1396) on the library's host range
coder (`masic_range_*`, csrc/cdf.cu).  Same call surface — RangeEncoder(path).encode(symbols, cum_freq) / .close(),
RangeDecoder(path).decode(n, cum_freq) / .close(), prob_to_cum_freq — the byte format is this library's own (the PyPI
package is neither vendored by the reference nor installable offline, so its bytes cannot be pinned)."""
import ctypes as C

import numpy as np

from masic_b200 import _lib


class RangeEncoder:
    def __init__(self, filepath):
        self._path, self._iv = filepath, []

    def encode(self, data, cumFreq):
        cf = cumFreq
        for s in data:
            s = int(s)
            self._iv.append((int(cf[s]), int(cf[s + 1]) - int(cf[s]), int(cf[-1])))

    def close(self):
        lib = _lib.load()
        iv = np.ascontiguousarray(self._iv, dtype=np.int32).reshape(-1, 3)
        out = np.empty(8 * len(iv) + 16, dtype=np.uint8)
        n = C.c_int64()
        _lib.check(lib.masic_range_encode(iv.ctypes.data, len(iv), out.ctypes.data, out.size, C.byref(n)),
                   "masic_range_encode")
        with open(self._path, "wb") as f:
            f.write(out[:n.value].tobytes())
        self._iv = []


class RangeDecoder:
    def __init__(self, filepath):
        with open(filepath, "rb") as f:
            self._data = np.frombuffer(f.read(), dtype=np.uint8).copy()
        self._lib = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self._lib.masic_range_decoder_create(self._data.ctypes.data, self._data.size, C.byref(self._h)),
                   "masic_range_decoder_create")

    def decode(self, size, cumFreq):
        row = np.ascontiguousarray(cumFreq, dtype=np.int32)
        rows = np.ascontiguousarray(np.broadcast_to(row, (size, row.size)))
        out = np.empty(size, dtype=np.int32)
        for i in range(size):       # one symbol per call of the row decoder: each symbol sees the updated state
            _lib.check(self._lib.masic_range_decode_rows(self._h, rows[i].ctypes.data, 1, row.size,
                                                         out[i:].ctypes.data), "masic_range_decode_rows")
        return out.tolist()

    def close(self):
        if self._h:
            self._lib.masic_range_decoder_destroy(self._h)
            self._h = None

    __del__ = close


def prob_to_cum_freq(prob, resolution=1024):
    """Integer frequencies summing to `resolution` that greedily minimise the KL divergence to `prob`."""
    prob = np.asarray(prob, dtype=np.float64)
    freq = np.zeros(prob.size, dtype=np.int64)
    with np.errstate(divide="ignore", invalid="ignore"):
        for _ in range(resolution):
            freq[np.nanargmax(prob / freq)] += 1
    return [0] + np.cumsum(freq).tolist()
