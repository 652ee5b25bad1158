"""compressai.ans — the reference's rANS extension, kept verbatim (north-star): re-export the binary
compiled from the reference's own sources (`make -C oracle ref`) or an installed one."""
from masic_b200.entropy_models import _load_ans as _load

_m = _load()
BufferedRansEncoder = _m.BufferedRansEncoder
RansEncoder = _m.RansEncoder
RansDecoder = _m.RansDecoder
