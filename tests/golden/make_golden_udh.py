"""Golden fixture of the udh homography front-end from the UNMODIFIED reference (coremasic/mywork/model.py `Net`,
test2_real.py `h_adjust` restated verbatim in the chain below because test2_real.py cannot be imported: lpips,
pytorch_msssim, imageio are absent) -> tests/golden/udh.npz.

    python tests/golden/make_golden_udh.py"""
import hashlib
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
from oracle import refimport, udh as OU  # noqa: E402


def main():
    refimport.import_masic()                      # puts the reference's mywork dir and the shims on sys.path
    import kornia
    import model as ref_model                     # coremasic/mywork/model.py, unmodified
    assert ref_model.__file__.startswith("/root/reference/"), ref_model.__file__
    torch.manual_seed(0)
    net = ref_model.Net(patch_size=128).eval()
    sd = net.state_dict()
    a, b, corners = OU.synthetic_patches(2, seed=3)
    with torch.no_grad():
        delta = net(a, b)
        c = corners - corners[:, 0].view(-1, 1, 2)                       # test2_real.py:203
        h = kornia.get_perspective_transform(c, c + delta)               # :207
        h = torch.inverse(h)                                             # :208
        # h_adjust(d1.shape[-2], d1.shape[-1], pic_size, pic_size, h_matrix), test2_real.py:54-64, :209
        A, B = 1216 / 256, 2176 / 256
        h[:, 0, :] = A * h[:, 0, :]
        h[:, :, 0] = (1. / A) * h[:, :, 0]
        h[:, 1, :] = B * h[:, 1, :]
        h[:, :, 1] = (1. / B) * h[:, :, 1]
        h_get = net.get_h(a, b, corners)                                 # model.py:102-111
    out = {"a": a.numpy(), "b": b.numpy(), "corners": corners.numpy(), "delta": delta.numpy(), "h_1216x2176": h.numpy(),
           "h_get_h": h_get.numpy(), "keys": np.array(list(sd.keys())),
           "shapes": np.array([str(tuple(v.shape)) for v in sd.values()]),
           "sha_fc5": np.array(hashlib.sha256(sd["fc.5.weight"].numpy().tobytes()).hexdigest()),
           "sha_cnn0": np.array(hashlib.sha256(sd["cnn.0.layers.0.weight"].numpy().tobytes()).hexdigest())}
    np.savez_compressed(HERE / "udh.npz", **out)
    print("wrote", HERE / "udh.npz", "delta", delta[0].flatten()[:4], "h", h[0])


if __name__ == "__main__":
    main()
