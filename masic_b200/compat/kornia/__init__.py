"""`import kornia` shim: the two entry points MASIC uses, on the masic_b200 CUDA warp kernel."""
from masic_b200.kornia_compat import __version__, get_perspective_transform, warp_perspective  # noqa: F401
