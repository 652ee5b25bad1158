from masic_b200.layers import LowerBound, NonNegativeParametrizer  # noqa: F401

__all__ = ["LowerBound", "NonNegativeParametrizer"]
