"""Import-time stub for the PyPI `range_coder` package (absent offline; TEST INFRASTRUCTURE).
Only HSIC.compress/decompress touch it (MASIC.py:958,1221), which the oracle never calls."""


class _Unavailable:
    def __init__(self, *a, **k):
        raise RuntimeError("range_coder is not available in this environment")


RangeEncoder = RangeDecoder = _Unavailable


def prob_to_cum_freq(*a, **k):
    raise RuntimeError("range_coder is not available in this environment")
