"""kornia.warp_perspective / get_perspective_transform as used by MASIC (kornia 0.5.0 semantics,
call sites coremasic/mywork/MASIC.py:638,644,781,821,833; test2_real.py:209) on the CUDA warp kernel."""
from __future__ import annotations

import torch

from . import torch_ops  # noqa: F401  (registers torch.ops.masic_b200.*)

__version__ = "0.5.0-masic_b200"


def warp_perspective(src, M, dsize, mode="bilinear", padding_mode="zeros", align_corners=None):
    """bilinear / zeros / align_corners=True (the 0.5.0 default) on the sm_100a warp kernel."""
    if mode != "bilinear" or padding_mode != "zeros" or align_corners is False:
        raise NotImplementedError("masic_b200 implements the configuration MASIC uses: bilinear, zeros, "
                                  "align_corners=True")
    n, c, h, w = src.shape
    if c <= 8:
        return torch.ops.masic_b200.warp_perspective(src, M, int(dsize[0]), int(dsize[1]))
    # wide feature maps (the CQE net warps 32-channel maps, MASIC.py:1479-1480): 8 channels per launch
    outs = [torch.ops.masic_b200.warp_perspective(src[:, i:i + 8].contiguous(), M, int(dsize[0]), int(dsize[1]))
            for i in range(0, c, 8)]
    return torch.cat(outs, dim=1)


def get_perspective_transform(src, dst):
    """4-point DLT (runs once per pair, upstream of the hot path: plain torch.linalg.solve)."""
    b = src.shape[0]
    rows = []
    for i in range(4):
        x, y = src[:, i, 0], src[:, i, 1]
        u, v = dst[:, i, 0], dst[:, i, 1]
        o, z = torch.ones_like(x), torch.zeros_like(x)
        rows.append(torch.stack([x, y, o, z, z, z, -x * u, -y * u], dim=1))
        rows.append(torch.stack([z, z, z, x, y, o, -x * v, -y * v], dim=1))
    A = torch.stack(rows, dim=1)
    rhs = dst.reshape(b, 8, 1)
    X = torch.linalg.solve(A, rhs)
    M = torch.ones(b, 9, device=src.device, dtype=src.dtype)
    M[:, :8] = X[..., 0]
    return M.view(b, 3, 3)
