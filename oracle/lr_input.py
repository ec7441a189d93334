"""CPU restatement of the training-sample assembly right before the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows climsr/data/sr/climate_dataset.py:152-172 (_get_training_sample: np.flipud, np.fliplr, np.rot90(k) applied to the HR
raster, the elevation and the land mask, then the LR raster = A.Resize(hr/4, INTER_NEAREST) of the augmented HR),
:127-141 (_common_to_tensor: elevation_lr by the same resize) and :98-121 (_concat_if_needed: x = [lr, elevation_lr,
mask_lr], mask_lr the same resize of the float mask).  albumentations' Resize is cv2.resize(img, (w, h), INTER_NEAREST)
(third-party, not vendored): for an integer factor it samples src[floor(y * s), floor(x * s)], i.e. the TOP-LEFT pixel
of every s x s block - pinned against cv2 itself in tests/golden/lr_input.npz (oracle/make_golden.py).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np


def augment(img: np.ndarray, v_flip: bool, h_flip: bool, rot_k: int) -> np.ndarray:
    """climate_dataset.py:152-170, in the reference's order."""
    if v_flip:
        img = np.flipud(img)
    if h_flip:
        img = np.fliplr(img)
    if rot_k:
        img = np.rot90(img, rot_k)
    return np.ascontiguousarray(img)


def resize_nearest(img: np.ndarray, scale: int = 4) -> np.ndarray:
    """cv2.resize(img, (W // scale, H // scale), interpolation=cv2.INTER_NEAREST) for H, W multiples of scale."""
    return np.ascontiguousarray(img[::scale, ::scale])


def training_sample(hr: np.ndarray, elev: np.ndarray, mask: np.ndarray, code: int, scale: int = 4):
    """One sample: code bit0 = vertical flip, bit1 = horizontal flip, bits 2-3 = rot90 factor.
    Returns x (3,h,w) = [lr, elevation_lr, mask_lr] and the augmented hr, elev, mask (H,W)."""
    v, hf, k = bool(code & 1), bool(code & 2), (code >> 2) & 3
    hr2, el2, mk2 = (augment(a, v, hf, k) for a in (hr, elev, mask))
    x = np.stack([resize_nearest(hr2, scale), resize_nearest(el2, scale), resize_nearest(mk2.astype(np.float32), scale)]).astype(np.float32)
    return x, hr2, el2, mk2


def training_batch(hr: np.ndarray, elev: np.ndarray, mask: np.ndarray, codes: Optional[Sequence[int]], scale: int = 4) -> Tuple[np.ndarray, ...]:
    """(N,1,H,W) tensors -> x (N,3,h,w), hr', elev', mask' (N,1,H,W)."""
    n = hr.shape[0]
    outs = [training_sample(hr[i, 0], elev[i, 0], mask[i, 0], int(codes[i]) if codes is not None else 0, scale) for i in range(n)]
    return (np.stack([o[0] for o in outs]),) + tuple(np.stack([o[j] for o in outs])[:, None] for j in (1, 2, 3))
