"""On-device mirror of the reference's ``MinMaxScaler`` for the inference path.

Reference: ``climsr/data/normalization.py:25-84``.  The reference normalises every LR raster with numpy inside the dataset
(``geo_tiff_inference_dataset.py:161-166``), assembles ``[raster, elevation_lr, mask_lr]`` (``:101-121``), and after the
forward calls ``.cpu().numpy()``, ``scaler.denormalize(arr, min, max)`` and ``arr[~mask] = nan`` per raster
(``climsr/inference/inference.py:70-76``).  Here both directions are one CUDA pass over the batch (float64 arithmetic, one
rounding to float32 - bit-identical to the reference under NumPy >= 2, see ``tests/golden/normalization.npz``), so rasters
stay on the GPU from the raw LR values to the denormalised, land-masked result.  No CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import torch
from torch import Tensor

from ._lib import CsrError, check, current_stream_ptr, lib

Stats = Union[Tensor, Sequence[float], float]


def _stats(v: Stats, n: int, dev) -> Tensor:
    t = torch.as_tensor(v, dtype=torch.float64).reshape(-1)
    if t.numel() == 1 and n > 1:
        t = t.expand(n)
    if t.numel() != n:
        raise ValueError(f"expected {n} min/max values, got {t.numel()}")
    return t.contiguous().to(dev)


class MinMaxScaler:
    """Same constructor and method names as ``climsr.data.normalization.MinMaxScaler`` (``:25-35``); operates on CUDA tensors."""

    def __init__(self, eps: Optional[float] = 1e-8, feature_range: Optional[Tuple[float, float]] = (0.0, 1.0),
                 nan_substitution: Optional[float] = 0.0):
        self.eps = eps
        self.feature_range = feature_range
        self.nan_substitution = nan_substitution
        self.a, self.b = self.feature_range

    def normalize(self, arr: Tensor, min: Stats, max: Stats, extras: Sequence[Tensor] = ()) -> Tensor:
        """``arr`` (N,h,w) or (N,1,h,w) raw rasters (NaN = missing) -> (N, 1+len(extras), h, w) fp32: channel 0 the normalised
        raster with NaN -> ``nan_substitution`` (``_normalize``, ``:37-61``), then the shared (h,w) planes of ``extras``
        (elevation_lr, mask_lr) as in ``geo_tiff_inference_dataset.py:101-121``."""
        if not arr.is_cuda:
            raise CsrError("climsr_b200.normalization runs on CUDA (sm_100a) only; there is no CPU fallback")
        if arr.dim() == 4:
            if arr.shape[1] != 1:
                raise ValueError("normalize expects single-channel rasters")
            arr = arr[:, 0]
        if arr.dim() != 3:
            raise ValueError(f"normalize expects (N,h,w) or (N,1,h,w), got {tuple(arr.shape)}")
        if len(extras) > 2:
            raise ValueError("at most two extra planes (elevation_lr, mask_lr)")
        n, h, w = arr.shape
        raw = arr.contiguous().float()
        ex = []
        for e in extras:
            e = e.to(arr.device).float().reshape(-1)
            if e.numel() != h * w:
                raise ValueError("extra planes must be (h,w)")
            ex.append(e.contiguous())
        mn, mx = _stats(min, n, arr.device), _stats(max, n, arr.device)
        out = torch.empty((n, 1 + len(ex), h, w), dtype=torch.float32, device=arr.device)
        with torch.cuda.device(arr.device):
            check(lib.csr_minmax_normalize(raw.data_ptr(), n, h, w, mn.data_ptr(), mx.data_ptr(), float(self.a), float(self.b), float(self.eps),
                                           float(self.nan_substitution), ex[0].data_ptr() if ex else None,
                                           ex[1].data_ptr() if len(ex) > 1 else None, out.data_ptr(), current_stream_ptr()),
                  "csr_minmax_normalize")
        return out

    def denormalize(self, arr: Tensor, min: Stats, max: Stats, mask: Optional[Tensor] = None) -> Tensor:
        """``arr`` (N,1,H,W) network output -> (N,1,H,W) fp32 ``(arr - min_) / scale`` (``_denormalize``, ``:63-84``); with
        ``mask`` ((1,1,H,W) shared or (N,1,H,W); > 0 = land) the pixels outside it become NaN (``inference.py:75``)."""
        if not arr.is_cuda:
            raise CsrError("climsr_b200.normalization runs on CUDA (sm_100a) only; there is no CPU fallback")
        if arr.dim() != 4 or arr.shape[1] != 1:
            raise ValueError(f"denormalize expects (N,1,H,W), got {tuple(arr.shape)}")
        n, _, h, w = arr.shape
        sr = arr.contiguous().float()
        if mask is None:
            m = torch.ones((1, 1, h, w), dtype=torch.float32, device=arr.device)
        else:
            m = mask.to(arr.device).float().contiguous()
            if m.numel() not in (h * w, n * h * w):
                raise ValueError(f"mask {tuple(mask.shape)} does not match {tuple(arr.shape)}")
        mn, mx = _stats(min, n, arr.device), _stats(max, n, arr.device)
        out = torch.empty_like(sr)
        with torch.cuda.device(arr.device):
            check(lib.csr_minmax_denormalize_mask(sr.data_ptr(), m.data_ptr(), 1 if (m.numel() == n * h * w and n > 1) else 0, n, h, w,
                                                  mn.data_ptr(), mx.data_ptr(), float(self.a), float(self.b), float(self.eps), out.data_ptr(),
                                                  current_stream_ptr()), "csr_minmax_denormalize_mask")
        return out
