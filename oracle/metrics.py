"""CPU restatement of the masked validation/test step and its 16 metrics.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows climsr/core/task.py:262-300 (common_val_test_step: denormalise, zero-fill
where mask==0, loss) and :302-380 (metric set + routing), climsr/data/normalization.py:
63-84 (MinMaxScaler._denormalize) / :115 (StandardScaler._denormalize) and
climsr/metrics/regression_accuracy.py:15-22.

torchmetrics (PSNR, SSIM, MAE, MSE, MAPE, SMAPE, R2Score) is a third-party
dependency that is NOT vendored in /root/reference and is unpinned there
(arrives with unpinned pytorch-lightning, environment.yml:23; the class names
imported at core/task.py:13-21 bound it to 0.5 <= v < 0.8).  Its published
formulas for that API are restated below -> PARITY UNPINNED for those seven;
RegressionAccuracy is pinned by the reference's nine known-answer cases
(tests/metrics/test_regresion_accuracy.py:12-108), replayed in tests/.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

ACC_EPS = (0.1, 0.25, 0.5, 0.75, 1.0, 1.25, 1.5, 2.0)          # core/task.py:303-310
ACC_KEYS = ("acc@0.1", "acc@0.25", "acc@0.5", "acc@0.75", "acc@1", "acc@01.25", "acc@1.5", "acc@2")  # :318-326 (typo kept)
MAPE_EPS = 1.17e-06


def minmax_denormalize(arr: Tensor, mn: Tensor, mx: Tensor, a: float = -1.0, b: float = 1.0, eps: float = 1e-8) -> Tensor:
    """normalization.py:63-84 for tensors: per-sample broadcast via permute(1,2,3,0)."""
    scale = (b - a) / ((mx - mn) + eps)
    min_ = a - mn * scale
    out = arr.permute(1, 2, 3, 0) - min_
    out = out / scale
    return out.permute(3, 0, 1, 2)


def zscore_denormalize(arr: Tensor, mean: float, std: float) -> Tensor:
    """normalization.py:115."""
    return arr * std + mean


def regression_accuracy(preds: Tensor, target: Tensor, eps: float) -> Tensor:
    """regression_accuracy.py:15-22: sum(|p-t| <= eps) / numel (float division of int counters)."""
    assert preds.shape == target.shape
    correct = torch.sum(torch.abs(preds - target) <= eps)
    return correct.float() / target.numel()


def psnr(preds: Tensor, target: Tensor) -> Tensor:
    """torchmetrics PSNR(data_range=None, base=10, dim=None): range from target min/max with 0.0 initial states."""
    mn = min(float(target.min()), 0.0)
    mx = max(float(target.max()), 0.0)
    mse = torch.mean((preds - target) ** 2)
    return 10.0 * torch.log10(torch.tensor((mx - mn) ** 2, dtype=mse.dtype) / mse)


def _gauss1d(k: int, sigma: float, dtype) -> Tensor:
    dist = torch.arange((1 - k) / 2, (1 + k) / 2, 1, dtype=dtype)
    g = torch.exp(-((dist / sigma) ** 2) / 2)
    return g / g.sum()


def ssim(preds: Tensor, target: Tensor, k: int = 11, sigma: float = 1.5, k1: float = 0.01, k2: float = 0.03) -> Tensor:
    """torchmetrics SSIM() defaults (functional _ssim_compute of 0.5-0.7)."""
    data_range = max(float(preds.max() - preds.min()), float(target.max() - target.min()))
    c1 = (k1 * data_range) ** 2
    c2 = (k2 * data_range) ** 2
    c = preds.shape[1]
    g = _gauss1d(k, sigma, preds.dtype)
    kernel = torch.outer(g, g).expand(c, 1, k, k)
    p = k // 2
    pp = F.pad(preds, (p, p, p, p), mode="reflect")
    tp = F.pad(target, (p, p, p, p), mode="reflect")
    stack = torch.cat((pp, tp, pp * pp, tp * tp, pp * tp))
    out = F.conv2d(stack, kernel, groups=c).split(preds.shape[0])
    mu_p2, mu_t2, mu_pt = out[0] ** 2, out[1] ** 2, out[0] * out[1]
    s_p2, s_t2, s_pt = out[2] - mu_p2, out[3] - mu_t2, out[4] - mu_pt
    idx = ((2 * mu_pt + c1) * (2 * s_pt + c2)) / ((mu_p2 + mu_t2 + c1) * (s_p2 + s_t2 + c2))
    idx = idx[..., p:-p, p:-p]
    return idx.mean()


def mape(preds: Tensor, target: Tensor) -> Tensor:
    return torch.mean(torch.abs(preds - target) / torch.clamp(torch.abs(target), min=MAPE_EPS))


def smape(preds: Tensor, target: Tensor) -> Tensor:
    return torch.mean(2 * torch.abs(preds - target) / torch.clamp(torch.abs(target) + torch.abs(preds), min=MAPE_EPS))


def r2(preds: Tensor, target: Tensor) -> Tensor:
    n = target.numel()
    rss = torch.sum((target - preds) ** 2)
    tss = torch.sum(target * target) - torch.sum(target) ** 2 / n
    return 1 - rss / tss


def compute_metrics(norm_sr: Tensor, norm_hr: Tensor, den_sr: Tensor, den_hr: Tensor) -> Dict[str, Tensor]:
    """core/task.py:342-380 routing: ssim,mape <- normalised; r2 <- flattened denormalised; rest <- denormalised."""
    res: Dict[str, Tensor] = {}
    for key, eps in zip(ACC_KEYS, ACC_EPS):
        res[key] = regression_accuracy(den_sr, den_hr, eps)
    res["psnr"] = psnr(den_sr, den_hr)
    res["ssim"] = ssim(norm_sr, norm_hr)
    res["mae"] = torch.mean(torch.abs(den_sr - den_hr))
    res["mse"] = torch.mean((den_sr - den_hr) ** 2)
    res["rmse"] = torch.sqrt(res["mse"])
    res["mape"] = mape(norm_sr, norm_hr)
    res["smape"] = smape(den_sr, den_hr)
    res["r2"] = r2(den_sr.flatten(), den_hr.flatten())
    return res


def val_test_step(sr: Tensor, hr: Tensor, original: Tensor, mask: Tensor,
                  mn: Optional[Tensor] = None, mx: Optional[Tensor] = None,
                  zscore: Optional[Tuple[float, float]] = None, loss: str = "l1",
                  feature_range: Tuple[float, float] = (-1.0, 1.0), dtype=torch.float64) -> Dict[str, Tensor]:
    """common_val_test_step arithmetic, core/task.py:262-300, on explicit tensors.

    sr/hr/original/mask: (N,1,H,W); mn/mx: (N,) for min-max; zscore=(mean,std) otherwise.
    Returns the metric dict plus 'loss' (== 'normalized_loss', task.py:293-294).
    """
    sr = sr.detach().to(dtype).clone()
    hr = hr.detach().to(dtype).clone()
    original = original.detach().to(dtype).clone()
    if zscore is not None:
        den_sr = zscore_denormalize(sr, zscore[0], zscore[1])                 # task.py:283
    else:
        den_sr = minmax_denormalize(sr, mn.to(dtype), mx.to(dtype), *feature_range)  # task.py:285
    den_sr = den_sr.clone()
    ocean = ~mask.bool()
    sr[ocean] = 0.0                                                           # task.py:288-291
    hr[ocean] = 0.0
    den_sr[ocean] = 0.0
    original[ocean] = 0.0
    lv = F.l1_loss(sr, hr) if loss == "l1" else F.mse_loss(sr, hr)            # task.py:141,293-294
    out = compute_metrics(sr, hr, den_sr, original)
    out["loss"] = lv
    return out


def denormalized_original(hr: Tensor, mn: Tensor, mx: Tensor, feature_range=(-1.0, 1.0)) -> Tensor:
    """original = denormalize(hr) for synthetic targets."""
    return minmax_denormalize(hr, mn, mx, *feature_range).contiguous()


def psnr_from_sums(sse: float, n: int, tmin: float, tmax: float) -> float:
    r = max(tmax, 0.0) - min(tmin, 0.0)
    return 10.0 * math.log10(r * r / (sse / n))
