"""Seeded synthetic weights and inputs shared by the oracle, tests and bench.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Shapes/names follow the reference generator's ``state_dict``
(climsr/models/esrgan.py:22-26,72-87 and climsr/models/srcnn.py:9-11); the
distributions follow nn.Conv2d's default init (kaiming-uniform a=sqrt(5) ==
U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias), which is what the
reference uses (its 0.1-scaled ESRGAN init is commented out, esrgan.py:30).
Input recipes follow SURVEY.md section 8d.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch


def conv_specs(in_channels: int = 3, out_channels: int = 1, nf: int = 64, nb: int = 23, gc: int = 32,
               scaling_factor: int = 4):
    """Ordered list of (name, cout, cin, kh, kw) for every conv of the generator.

    Order == reference ``state_dict`` order (esrgan.py:72-87).
    """
    specs = [("conv_first", nf, in_channels, 3, 3)]
    for i in range(nb):
        for r in (1, 2, 3):
            for k in range(1, 5):
                specs.append((f"RRDB_trunk.{i}.RDB{r}.conv{k}", gc, nf + (k - 1) * gc, 3, 3))
            specs.append((f"RRDB_trunk.{i}.RDB{r}.conv5", nf, nf + 4 * gc, 3, 3))
    specs.append(("trunk_conv", nf, nf, 3, 3))
    specs.append(("upconv1", nf, nf, 3, 3))
    if scaling_factor == 4:
        specs.append(("upconv2", nf, nf, 3, 3))
    specs.append(("HRconv", nf, nf, 3, 3))
    specs.append(("conv_last", out_channels, nf, 3, 3))
    specs.append(("srcnn.conv1", 64, 3, 9, 9))
    specs.append(("srcnn.conv2", 32, 64, 1, 1))
    specs.append(("srcnn.conv3", out_channels, 32, 5, 5))
    return specs


def make_state_dict(in_channels: int = 3, out_channels: int = 1, nf: int = 64, nb: int = 23, gc: int = 32,
                    scaling_factor: int = 4, seed: int = 0, gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic fp32 state_dict with the reference's names and shapes.

    ``gain`` multiplies every conv weight (not bias); gain > 1 gives the
    "trained-like" O(1) output range SURVEY.md section 8d asks for.
    """
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, co, ci, kh, kw in conv_specs(in_channels, out_channels, nf, nb, gc, scaling_factor):
        bound = 1.0 / math.sqrt(ci * kh * kw)
        w = (torch.rand((co, ci, kh, kw), generator=g, dtype=torch.float32) * 2 - 1) * bound * gain
        b = (torch.rand((co,), generator=g, dtype=torch.float32) * 2 - 1) * bound
        sd[name + ".weight"] = w
        sd[name + ".bias"] = b
    return sd


def make_inputs(n: int, in_channels: int, h: int, w: int, scale: int = 4, seed: int = 1,
                blocky_mask: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(x, elev, mask) as the reference dataset would hand them over.

    x ~ U(-1,1) (N,Cin,h,w); mask = (U(0,1) > 0.3) float (N,1,H,W); elev ~ U(-1,1)
    zeroed over the ocean; channel 2 of x (when present) is the nearest-/4 LR
    mask, as climate_dataset.py:107-118 concatenates [lr_var, elev_lr, mask_lr].
    """
    H, W = h * scale, w * scale
    gx = torch.Generator().manual_seed(seed)
    ge = torch.Generator().manual_seed(seed + 1)
    gm = torch.Generator().manual_seed(seed + 2)
    x = torch.rand((n, in_channels, h, w), generator=gx) * 2 - 1
    if blocky_mask:
        low = torch.rand((n, 1, max(H // 16, 1), max(W // 16, 1)), generator=gm)
        mask = torch.nn.functional.interpolate(low, size=(H, W), mode="nearest")
        mask = (mask > 0.3).float()
    else:
        mask = (torch.rand((n, 1, H, W), generator=gm) > 0.3).float()
    elev = (torch.rand((n, 1, H, W), generator=ge) * 2 - 1) * mask
    if in_channels >= 3:
        x[:, 2:3] = mask[:, :, ::scale, ::scale]
    return x.contiguous(), elev.contiguous(), mask.contiguous()


def make_targets(sr: torch.Tensor, seed: int = 4) -> Dict[str, torch.Tensor]:
    """hr / min / max / original for the masked val step."""
    n = sr.shape[0]
    g = torch.Generator().manual_seed(seed)
    hr = (sr.float().cpu() + 0.05 * torch.randn(sr.shape, generator=g)).clamp(-1, 1)
    g2 = torch.Generator().manual_seed(seed + 1)
    mn = -60 + 40 * torch.rand((n,), generator=g2)
    mx = 20 + 30 * torch.rand((n,), generator=g2)
    return {"hr": hr, "min": mn, "max": mx}


def make_discriminator_state_dict(seed: int = 0, in_channels: int = 1, width: int = 64, stages: int = 4) -> "OrderedDict[str, torch.Tensor]":
    """Seeded weights with the names / shapes / order of the reference discriminator's ``state_dict``
    (climsr/models/discriminator.py:8-40: nn.Sequential indices 1, 3, 5 (+7 per stage) for conv, BatchNorm, conv; then the
    two valid convs; ``classification.{0,1}``).  Conv / Linear: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like the default init;
    BatchNorm weight ~ U(0.5, 1.5), bias ~ U(-0.2, 0.2) so the affine part is exercised; fresh running statistics."""
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def uni(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    def conv(name, cout, cin, k=3):
        b = 1.0 / math.sqrt(cin * k * k)
        sd[name + ".weight"] = uni((cout, cin, k, k), b)
        sd[name + ".bias"] = uni((cout,), b)

    idx, cin, c = 0, in_channels, width
    for _ in range(stages):
        conv(f"feature_extraction.{idx + 1}", c, cin)
        bn = f"feature_extraction.{idx + 3}"
        sd[bn + ".weight"] = torch.rand((c,), generator=g) + 0.5
        sd[bn + ".bias"] = uni((c,), 0.2)
        sd[bn + ".running_mean"] = torch.zeros(c)
        sd[bn + ".running_var"] = torch.ones(c)
        sd[bn + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        conv(f"feature_extraction.{idx + 5}", c, c)
        cin, c, idx = c, c * 2, idx + 7
    conv(f"feature_extraction.{idx}", cin, cin)
    conv(f"feature_extraction.{idx + 2}", cin, cin)
    for name, o, i in (("classification.0", 100, 8192), ("classification.1", 1, 100)):
        b = 1.0 / math.sqrt(i)
        sd[name + ".weight"] = uni((o, i), b)
        sd[name + ".bias"] = uni((o,), b)
    return sd
