"""compressai.datasets (MASIC.py:32, test2_real.py:27): the stereo ImageFolder of masic_b200/datasets.py."""
from masic_b200.datasets import ImageFolder  # noqa: F401

__all__ = ["ImageFolder"]
