"""compressai.models: MASIC.py:34 imports CompressionModel and immediately shadows it (MASIC.py:40)."""
import torch.nn as nn

from . import utils  # noqa: F401


class CompressionModel(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("use masic_b200.hsic.HSIC; the upstream model zoo is outside the MASIC hot path")
