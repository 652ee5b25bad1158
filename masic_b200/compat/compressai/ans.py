"""compressai.ans — the rANS coder surface (RansEncoder / BufferedRansEncoder / RansDecoder) on the library's
native coder (masic_b200/rans.py, csrc/rans.cpp): same list API and byte format as the reference's pybind11
extension (compressai/cpp_exts/rans/rans_interface.cpp); $MASIC_ANS_MODULE substitutes another module."""
from masic_b200.entropy_models import _load_ans as _load

_m = _load()
BufferedRansEncoder = _m.BufferedRansEncoder
RansEncoder = _m.RansEncoder
RansDecoder = _m.RansDecoder
