"""`import compressai` resolved to masic_b200 (the surface coremasic/mywork/MASIC.py:18-38 imports)."""
from compressai import datasets, entropy_models, layers, models, ops  # noqa: F401

__version__ = "1.2.0b3.masic_b200"
_entropy_coder = "ans"
_available_entropy_coders = [_entropy_coder]


def set_entropy_coder(entropy_coder):
    global _entropy_coder
    if entropy_coder not in _available_entropy_coders:
        raise ValueError(f'Invalid entropy coder "{entropy_coder}", choose from'
                         f'({", ".join(_available_entropy_coders)}).')
    _entropy_coder = entropy_coder


def get_entropy_coder():
    return _entropy_coder


def available_entropy_coders():
    return _available_entropy_coders
