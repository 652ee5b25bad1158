"""CPU restatement of the two kornia 0.5.0 entry points MASIC uses (TEST INFRASTRUCTURE).

kornia is a third-party dependency of the reference that is NOT vendored in its tree
(`readme.md:12` pins `kornia==0.5.0`) and is not installed here, so this file restates
the published 0.5.0 algorithm from its documented behaviour:

  warp_perspective(src, M, dsize):
      N(h, w)   = [[2/(w-1), 0, -1], [0, 2/(h-1), -1], [0, 0, 1]]         (normal_transform_pixel)
      T         = N(dst) @ M @ inv(N(src))                                 (normalize_homography)
      grid      = transform_points(inv(T), meshgrid of ((i/(n-1)) - 0.5) * 2)   (x first)
      homogeneous divide with  scale = where(|z| > 1e-8, 1/(z + 1e-8), 1)
      out       = F.grid_sample(src, grid, 'bilinear', 'zeros', align_corners=True)

Call sites in the reference: coremasic/mywork/MASIC.py:638,644,781,821,833 (codec) and
:1461-1480 (CQE).  The reference has no test that pins warp results, so parity for this
function is "unpinned": this restatement IS the specification the CUDA warp is checked
against (tolerance item, <= 1e-4 abs on [0,1] images).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

__version__ = "0.5.0-restated"


def _normal_transform_pixel(h: int, w: int, device, dtype) -> torch.Tensor:
    eps = 1e-14
    wd = eps if w == 1 else w - 1.0
    hd = eps if h == 1 else h - 1.0
    t = torch.tensor([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]], device=device, dtype=dtype)
    t[0, 0] = t[0, 0] * 2.0 / wd
    t[1, 1] = t[1, 1] * 2.0 / hd
    return t.unsqueeze(0)


def _inverse_cast(m: torch.Tensor) -> torch.Tensor:
    dt = m.dtype if m.dtype in (torch.float32, torch.float64) else torch.float32
    return torch.inverse(m.to(dt)).to(m.dtype)


def _meshgrid_normalized(h: int, w: int, device, dtype) -> torch.Tensor:
    xs = torch.linspace(0, w - 1, w, device=device, dtype=dtype)
    ys = torch.linspace(0, h - 1, h, device=device, dtype=dtype)
    xs = (xs / (w - 1) - 0.5) * 2
    ys = (ys / (h - 1) - 0.5) * 2
    gx, gy = torch.meshgrid(xs, ys, indexing="ij")          # (w, h)
    grid = torch.stack((gx, gy)).transpose(1, 2)            # 2 x h x w
    return grid.unsqueeze(0).permute(0, 2, 3, 1)            # 1 x h x w x 2


def _transform_points(trans: torch.Tensor, pts: torch.Tensor) -> torch.Tensor:
    shape = list(pts.shape)
    pts = pts.reshape(-1, pts.shape[-2], pts.shape[-1])
    trans = trans.reshape(-1, trans.shape[-2], trans.shape[-1])
    trans = torch.repeat_interleave(trans, repeats=pts.shape[0] // trans.shape[0], dim=0)
    pts_h = F.pad(pts, (0, 1), "constant", 1.0)
    out_h = torch.bmm(pts_h, trans.permute(0, 2, 1))
    z = out_h[..., -1:]
    scale = torch.where(torch.abs(z) > 1e-8, 1.0 / (z + 1e-8), torch.ones_like(z))
    out = scale * out_h[..., :-1]
    shape[-1] = out.shape[-1]
    return out.reshape(shape)


def normalized_src_from_dst(M: torch.Tensor, src_hw, dst_hw) -> torch.Tensor:
    """inv(N_dst @ M @ inv(N_src)) — the matrix applied to the normalised destination grid."""
    h, w = src_hw
    ho, wo = dst_hw
    n_src = _normal_transform_pixel(h, w, M.device, M.dtype)
    n_dst = _normal_transform_pixel(ho, wo, M.device, M.dtype)
    t = n_dst @ (M @ _inverse_cast(n_src))
    return _inverse_cast(t)


def warp_perspective(src, M, dsize, mode="bilinear", padding_mode="zeros", align_corners=None):
    if align_corners is None:
        align_corners = True                      # 0.5.0 default (with a deprecation warning)
    b, _, h, w = src.shape
    ho, wo = dsize
    t_inv = normalized_src_from_dst(M, (h, w), (ho, wo))
    grid = _meshgrid_normalized(ho, wo, src.device, src.dtype).repeat(b, 1, 1, 1)
    grid = _transform_points(t_inv[:, None, None], grid)
    return F.grid_sample(src, grid, mode=mode, padding_mode=padding_mode, align_corners=align_corners)


def get_perspective_transform(src, dst):
    """4-point DLT (kornia 0.5.0: 8x8 system solved with torch.solve)."""
    b = src.shape[0]
    rows = []
    for i in range(4):
        x, y = src[:, i, 0], src[:, i, 1]
        u, v = dst[:, i, 0], dst[:, i, 1]
        o, z = torch.ones_like(x), torch.zeros_like(x)
        rows.append(torch.stack([x, y, o, z, z, z, -x * u, -y * u], dim=1))
        rows.append(torch.stack([z, z, z, x, y, o, -x * v, -y * v], dim=1))
    A = torch.stack(rows, dim=1)
    rhs = torch.stack([dst[:, 0, 0], dst[:, 0, 1], dst[:, 1, 0], dst[:, 1, 1],
                       dst[:, 2, 0], dst[:, 2, 1], dst[:, 3, 0], dst[:, 3, 1]], dim=1).unsqueeze(-1)
    X = torch.linalg.solve(A, rhs)
    M = torch.ones(b, 9, device=src.device, dtype=src.dtype)
    M[:, :8] = X[..., 0]
    return M.view(b, 3, 3)
