"""Stereo `ImageFolder` with the surface the reference's scripts use (compressai/datasets/utils.py:68-404, called at
test2_real.py:362-369, newtrain_codec_real.py:343-354 with `patch_size, split, transform, need_H, need_file_name,
root_add`).  I/O sits outside the hot path; this exists so that the scripts get past their dataset
lines on `compat/compressai`.  Layout: root/{train,test}/{left,right}/<same file names>.  One sample is

    (x1, x2, H, [file_name,] homo_patch1, homo_patch2, corners)

x1/x2: the (randomly cropped) RGB views through `transform`; H: the classical (SURF) homography when `need_H` —
needs an OpenCV build with xfeatures2d, which neither the reference pins nor this image has, so `need_H=True`
raises a clear error there and the scripts' `need_H=False` path returns the reference's 'None' placeholder;
homo patches: 128x128 grey crops of the 256x256-resized views, normalised with the mean ImageNet statistics, and
their corner coordinates — the inputs of the udh homography net (coremasic/mywork/model.py)."""
from __future__ import annotations

import random
from pathlib import Path

import numpy as np
import torch
from torch.utils.data import Dataset

MEAN = float(np.mean([0.485, 0.456, 0.406]))
STD = float(np.mean([0.229, 0.224, 0.225]))
HOMO_PIC, HOMO_PATCH, RHO = 256, 128, 45


def _read_rgb(path) -> np.ndarray:
    import cv2
    img = cv2.imread(str(path))
    if img is None:
        raise ValueError(f"cannot read image {path}")
    return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)


def classical_homography(im1: np.ndarray, im2: np.ndarray) -> torch.Tensor:
    """SURF + FLANN + RANSAC homography of the right view onto the left (datasets/utils.py:30-66)."""
    import cv2
    if not hasattr(cv2, "xfeatures2d"):
        raise RuntimeError("need_H=True needs OpenCV's xfeatures2d (SURF); use need_H=False and the udh homography "
                           "net, as the reference's test/train scripts do")
    surf = cv2.xfeatures2d.SURF_create(400)
    k1, d1 = surf.detectAndCompute(im1, None)
    k2, d2 = surf.detectAndCompute(im2, None)
    matches = cv2.FlannBasedMatcher(dict(algorithm=1, trees=5), dict(checks=50)).knnMatch(d1, d2, k=2)
    good = [m for m, n in matches if m.distance < 0.7 * n.distance]
    src = np.float32([k1[m.queryIdx].pt for m in good]).reshape(-1, 1, 2)
    dst = np.float32([k2[m.trainIdx].pt for m in good]).reshape(-1, 1, 2)
    Hm, _ = cv2.findHomography(src, dst, cv2.RANSAC, 5.0)
    return torch.from_numpy(Hm.astype(np.float32))


class ImageFolder(Dataset):
    def __init__(self, root, transform=None, patch_size=(256, 256), split="train", need_file_name=False,
                 root2="", need_root2=False, root_add="", need_H=True):
        if need_root2:
            raise NotImplementedError("root2 (quality-enhancement comparison set) is not part of the codec path")
        self.left_list, self.right_list = [], []
        for r in (root, root_add):
            if r == "":
                continue
            splitdir = Path(r) / split
            if not splitdir.is_dir():
                raise RuntimeError(f'Invalid directory "{r}"')
            self.left_list += sorted(str(p) for p in (splitdir / "left").glob("*"))
            self.right_list += sorted(str(p) for p in (splitdir / "right").glob("*"))
        self.patch_size, self.transform = tuple(patch_size), transform
        self.need_file_name, self.need_H = need_file_name, need_H

    def __len__(self):
        return len(self.left_list)

    @staticmethod
    def _homo_patch(img: np.ndarray) -> torch.Tensor:
        import cv2
        g = cv2.resize(img, (HOMO_PIC, HOMO_PIC))
        t = torch.from_numpy(np.ascontiguousarray(g)).permute(2, 0, 1).float().div(255)
        return ((t - MEAN) / STD).mean(dim=0, keepdim=True)

    def __getitem__(self, index):
        lp, rp = Path(self.left_list[index]), Path(self.right_list[index])
        if lp.name != rp.name:
            raise ValueError("cannot compare pictures.")
        img1, img2 = _read_rgb(lp), _read_rgb(rp)
        h, w, _ = img1.shape
        ph, pw = min(self.patch_size[0], h), min(self.patch_size[1], w)
        H = classical_homography(img1, img2) if self.need_H else "None"
        if ph == h:
            y0 = x0 = 0
        else:                                           # the reference's closed-interval randint, one short of the edge
            y0, x0 = random.randint(0, h - ph - 1), random.randint(0, w - pw - 1)
        img1, img2 = img1[y0:y0 + ph, x0:x0 + pw], img2[y0:y0 + ph, x0:x0 + pw]
        g1, g2 = self._homo_patch(img1), self._homo_patch(img2)
        if HOMO_PIC - RHO - HOMO_PATCH >= RHO:
            x = random.randint(RHO, HOMO_PIC - RHO - HOMO_PATCH)
            y = random.randint(RHO, HOMO_PIC - RHO - HOMO_PATCH)
        else:
            x = y = 0
        corners = torch.tensor([[x, y], [x + HOMO_PATCH, y], [x + HOMO_PATCH, y + HOMO_PATCH], [x, y + HOMO_PATCH]],
                               dtype=torch.float32)
        g1, g2 = g1[:, y:y + HOMO_PATCH, x:x + HOMO_PATCH], g2[:, y:y + HOMO_PATCH, x:x + HOMO_PATCH]
        a, b = (self.transform(img1), self.transform(img2)) if self.transform else (img1, img2)
        if self.need_file_name:
            return a, b, H, lp.name, g1, g2, corners
        return a, b, H, g1, g2, corners
