"""CPU restatement of the min-max scaler around the inference hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows climsr/data/normalization.py:37-61 (MinMaxScaler._normalize: missing indicator -> NaN, scale into
feature_range, NaN -> nan_substitution, float32) and :63-84 (MinMaxScaler._denormalize, numpy branch), the LR-input
assembly of climsr/data/sr/geo_tiff_inference_dataset.py:101-121,161-166 (channels [normalised raster, elevation_lr,
mask_lr]) and the post-processing of climsr/inference/inference.py:73-76 (denormalise with the raster's min/max, NaN
outside the land mask).  Arithmetic is float64 with one final rounding to float32, which is what the reference does under
NumPy >= 2 when min/max are float64 scalars (pandas) - pinned by tests/golden/normalization.npz, generated from the
reference's own class by oracle/make_golden.py.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np


def minmax_coeffs(mn: float, mx: float, feature_range: Tuple[float, float] = (-1.0, 1.0), eps: float = 1e-8):
    a, b = feature_range
    scale = (b - a) / ((np.float64(mx) - np.float64(mn)) + eps)          # normalization.py:53-55 / :70-72
    return scale, a - np.float64(mn) * scale


def normalize(arr: np.ndarray, mn: float, mx: float, feature_range=(-1.0, 1.0), eps: float = 1e-8, nan_substitution: float = 0.0,
              missing_indicator: Optional[float] = None) -> np.ndarray:
    """normalization.py:37-61 for given min / max."""
    out = arr.astype(np.float64)
    if missing_indicator:
        out[arr == missing_indicator] = np.nan                           # :45-46
    scale, min_ = minmax_coeffs(mn, mx, feature_range, eps)
    out = out * scale                                                    # :57
    out += min_                                                          # :58
    out[np.isnan(out)] = nan_substitution                                # :60
    return out.astype(np.float32)


def denormalize(arr: np.ndarray, mn: float, mx: float, feature_range=(-1.0, 1.0), eps: float = 1e-8) -> np.ndarray:
    """normalization.py:63-80 (numpy branch); float64 result like the reference."""
    scale, min_ = minmax_coeffs(mn, mx, feature_range, eps)
    return (arr.astype(np.float64) - min_) / scale


def lr_input(raw: np.ndarray, mins: Sequence[float], maxes: Sequence[float], elev_lr: Optional[np.ndarray], mask_lr: Optional[np.ndarray],
             feature_range=(-1.0, 1.0)) -> np.ndarray:
    """(N,h,w) raw rasters -> (N,C,h,w) generator input: geo_tiff_inference_dataset.py:161-166 then :101-121."""
    planes = []
    for i in range(raw.shape[0]):
        ch = [normalize(raw[i], mins[i], maxes[i], feature_range)]
        if elev_lr is not None:
            ch.append(elev_lr.astype(np.float32))
        if mask_lr is not None:
            ch.append(mask_lr.astype(np.float32))
        planes.append(np.stack(ch))
    return np.stack(planes)


def postprocess(sr: np.ndarray, mask: np.ndarray, mins: Sequence[float], maxes: Sequence[float], feature_range=(-1.0, 1.0)) -> np.ndarray:
    """(N,1,H,W) network output -> denormalised float32 rasters with NaN outside the land mask: inference.py:73-76."""
    out = np.empty(sr.shape, dtype=np.float32)
    for i in range(sr.shape[0]):
        arr = denormalize(sr[i, 0], mins[i], maxes[i], feature_range)
        m = mask[i, 0] if mask.shape[0] > 1 else mask[0, 0]
        arr[~(m > 0)] = np.nan
        out[i, 0] = arr.astype(np.float32)
    return out
