"""compressai._CXX — pmf_to_quantized_cdf on the library's native host code."""
from masic_b200.ops import pmf_to_quantized_cdf  # noqa: F401
