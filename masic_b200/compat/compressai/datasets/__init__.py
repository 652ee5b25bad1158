"""compressai.datasets: I/O, outside the hot path.  MASIC.py:32 only needs the name."""


class ImageFolder:
    def __init__(self, *a, **k):
        raise NotImplementedError("dataset loading is outside the masic_b200 hot path; benchmarks use synthetic pairs")
