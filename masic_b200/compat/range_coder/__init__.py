"""`range_coder` (PyPI, unavailable offline) is only touched inside HSIC.compress/decompress
(MASIC.py:958,1221); importing it must merely succeed."""


class _Unavailable:
    def __init__(self, *a, **k):
        raise RuntimeError("range_coder is not installed; the y-bitstream path of HSIC.compress is not built")


RangeEncoder = RangeDecoder = _Unavailable


def prob_to_cum_freq(*a, **k):
    raise RuntimeError("range_coder is not installed")
