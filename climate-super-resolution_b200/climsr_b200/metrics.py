"""Masked validation/test metrics on device (mirror of climsr/core/task.py:262-300 + compute_metrics :342-380)."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from ._lib import METRIC_KEYS, NUM_METRICS, check, current_stream_ptr, lib


def masked_val_metrics_raw(sr: Tensor, hr: Tensor, original: Tensor, mask: Tensor, min_vals: Optional[Tensor] = None,
                           max_vals: Optional[Tensor] = None, zscore: Optional[Tuple[float, float]] = None,
                           feature_range: Tuple[float, float] = (-1.0, 1.0), eps: float = 1e-8, scaler=None) -> Tensor:
    """Returns the CSR_NUM_METRICS fp32 vector on device (no host sync).

    ``scaler`` (a climsr_b200.normalization.MinMaxScaler or the reference's own, normalization.py:22-35) supplies
    ``feature_range`` and ``eps``; otherwise they default to the Hydra datamodule's (-1, 1) and the scaler's 1e-8.
    min_vals / max_vals are taken as float64, like the reference's batch["min"] / batch["max"]."""
    if scaler is not None:
        feature_range, eps = tuple(scaler.feature_range), float(scaler.eps)
    n, c, h, w = sr.shape
    if c != 1:
        raise ValueError("metrics expect single-channel (N,1,H,W) tensors")
    dev = sr.device
    f = lambda t: t.detach().to(dev).contiguous().float()  # noqa: E731
    sr, hr, original, mask = f(sr), f(hr), f(original), f(mask)
    f64 = lambda t: t.detach().to(dev).contiguous().double()  # noqa: E731
    mn = f64(min_vals) if min_vals is not None else None
    mx = f64(max_vals) if max_vals is not None else None
    if mn is not None and (mn.numel() != n or mx.numel() != n):
        raise ValueError(f"min_vals / max_vals must hold one value per sample ({n})")
    if (mn is None) != (mx is None) or (mn is None and zscore is None):
        raise ValueError("give min_vals and max_vals (min-max scaler) or zscore=(mean, std)")
    zm, zs = zscore if zscore is not None else (0.0, 1.0)
    out = torch.empty(NUM_METRICS, dtype=torch.float32, device=dev)
    nbytes = lib.csr_metrics_scratch_bytes(n, h, w)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.csr_masked_metrics(sr.data_ptr(), hr.data_ptr(), original.data_ptr(), mask.data_ptr(),
                                     mn.data_ptr() if mn is not None else None, mx.data_ptr() if mx is not None else None,
                                     zm, zs, float(feature_range[0]), float(feature_range[1]), float(eps), n, h, w, out.data_ptr(),
                                     scratch.data_ptr(), nbytes, current_stream_ptr()), "csr_masked_metrics")
    return out


def masked_val_metrics(sr, hr, original, mask, min_vals=None, max_vals=None, zscore=None, feature_range=(-1.0, 1.0),
                       prefix: str = "val", loss: str = "l1", eps: float = 1e-8, scaler=None) -> Dict[str, Tensor]:
    """Dict with the reference's metric keys (``{prefix}/acc@0.1`` ... ``{prefix}/r2``) plus loss / normalized_loss."""
    v = masked_val_metrics_raw(sr, hr, original, mask, min_vals, max_vals, zscore, feature_range, eps, scaler)
    out = {f"{prefix}/{k}": v[i] for i, k in enumerate(METRIC_KEYS[:16])}
    lv = v[16] if loss == "l1" else v[17]
    out[f"{prefix}/normalized_loss"] = lv
    out[f"{prefix}/loss"] = lv
    return out
