"""CPU fp32 restatement of the reference ESRGAN generator forward (+ autograd backward).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference lines it follows; paths are relative to /root/reference.

The arithmetic of the reference lives in PyTorch ATen (conv2d, leaky_relu,
cat, upsample_nearest2d; environment.yml:9 pins pytorch=1.10.0).  This file
restates the *graph* with plain ``torch.nn.functional`` calls on a
``state_dict`` (no nn.Module of the reference is imported), and
oracle/np_ops.py restates the *arithmetic* of those ops in numpy so the two
can be cross-checked without trusting either.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def _conv(sd: Dict[str, Tensor], name: str, x: Tensor, pad: int) -> Tensor:
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=1, padding=pad)


def _act(x: Tensor, slope: float, masks: Optional[Dict[str, Tensor]], name: str) -> Tensor:
    """LeakyReLU(slope) / ReLU (slope 0).  With ``masks`` the activation pattern is GIVEN (True = pass, False = * slope) instead of
    taken from the sign of x: the bf16-emulating form the backward parity tests use - the sm_100a path's forward runs with bf16
    activations, and the ~1 % of pre-activations whose sign that flips would otherwise dominate a gradient comparison."""
    if masks is None:
        return F.leaky_relu(x, slope) if slope else F.relu(x)
    m = masks[name]
    assert m.shape == x.shape, (name, tuple(m.shape), tuple(x.shape))
    return x * torch.where(m, torch.ones((), dtype=x.dtype), torch.full((), slope, dtype=x.dtype))


def rdb_forward(sd: Dict[str, Tensor], prefix: str, x: Tensor, masks: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """ResidualDenseBlock.forward, climsr/models/esrgan.py:32-38."""
    x1 = _act(_conv(sd, prefix + ".conv1", x, 1), 0.2, masks, prefix + ".conv1")
    x2 = _act(_conv(sd, prefix + ".conv2", torch.cat((x, x1), 1), 1), 0.2, masks, prefix + ".conv2")
    x3 = _act(_conv(sd, prefix + ".conv3", torch.cat((x, x1, x2), 1), 1), 0.2, masks, prefix + ".conv3")
    x4 = _act(_conv(sd, prefix + ".conv4", torch.cat((x, x1, x2, x3), 1), 1), 0.2, masks, prefix + ".conv4")
    x5 = _conv(sd, prefix + ".conv5", torch.cat((x, x1, x2, x3, x4), 1), 1)
    return x5 * 0.2 + x


def rrdb_forward(sd: Dict[str, Tensor], prefix: str, x: Tensor, masks: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """ResidualInResidualDenseBlock.forward, esrgan.py:50-54."""
    out = rdb_forward(sd, prefix + ".RDB1", x, masks)
    out = rdb_forward(sd, prefix + ".RDB2", out, masks)
    out = rdb_forward(sd, prefix + ".RDB3", out, masks)
    return out * 0.2 + x


def srcnn_forward(sd: Dict[str, Tensor], prefix: str, x: Tensor, masks: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """SRCNN.forward, climsr/models/srcnn.py:13-18 (9x9 p4, 1x1, 5x5 p2)."""
    out = _act(_conv(sd, prefix + ".conv1", x, 4), 0.0, masks, prefix + ".conv1")
    out = _act(_conv(sd, prefix + ".conv2", out, 0), 0.0, masks, prefix + ".conv2")
    return _conv(sd, prefix + ".conv3", out, 2)


def count_rrdb(sd: Dict[str, Tensor]) -> int:
    nb = 0
    while f"RRDB_trunk.{nb}.RDB1.conv1.weight" in sd:
        nb += 1
    return nb


def generator_forward(sd: Dict[str, Tensor], x: Tensor, elev: Tensor, mask: Tensor,
                      taps: Optional[Dict[str, Tensor]] = None, masks: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """ESRGANGenerator.forward, esrgan.py:89-102.

    ``taps`` (optional dict) receives named intermediates for layer-level parity; ``masks`` (optional dict: layer name ->
    bool tensor) fixes the activation patterns (see ``_act``).
    """
    nb = count_rrdb(sd)
    fea = _conv(sd, "conv_first", x, 1)                                     # esrgan.py:90
    t = fea
    for i in range(nb):                                                     # esrgan.py:91
        t = rrdb_forward(sd, f"RRDB_trunk.{i}", t, masks)
        if taps is not None and i == 0:
            taps["rrdb0"] = t
    trunk = _conv(sd, "trunk_conv", t, 1)
    fea = fea + trunk                                                       # esrgan.py:92
    if taps is not None:
        taps["fea"] = fea
    fea = _act(_conv(sd, "upconv1", F.interpolate(fea, scale_factor=2, mode="nearest"), 1), 0.2, masks, "upconv1")  # :94
    if "upconv2.weight" in sd:                                              # esrgan.py:96-97
        fea = _act(_conv(sd, "upconv2", F.interpolate(fea, scale_factor=2, mode="nearest"), 1), 0.2, masks, "upconv2")
    out = _conv(sd, "conv_last", _act(_conv(sd, "HRconv", fea, 1), 0.2, masks, "HRconv"), 1)  # esrgan.py:99
    if taps is not None:
        taps["conv_last"] = out
    out = srcnn_forward(sd, "srcnn", torch.cat([out, elev, mask], 1), masks)       # esrgan.py:100
    return out


def generator_forward_backward(sd: Dict[str, Tensor], x: Tensor, elev: Tensor, mask: Tensor, hr: Tensor,
                               loss: str = "l1", masks: Optional[Dict[str, Tensor]] = None) -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
    """Training-step arithmetic: L1 (esrgan) / MSE (srcnn) pixel loss, core/task.py:141 and
    task/pl_generator_pre_training.py:29-30, then autograd backward (a6 in SURVEY.md section 8a).

    Returns (sr, loss, grads-by-name).
    """
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    sr = generator_forward(params, x, elev, mask, masks=masks)
    lv = F.l1_loss(sr, hr) if loss == "l1" else F.mse_loss(sr, hr)
    grads = torch.autograd.grad(lv, list(params.values()))
    return sr.detach(), lv.detach(), {k: g for k, g in zip(params.keys(), grads)}


def flops_per_hr_pixel(in_channels: int, nf: int, nb: int, gc: int, out_channels: int = 1) -> float:
    """Algorithmic forward FLOPs per output pixel (SURVEY.md section 8a aggregate model, scale 4)."""
    rdb = gc * (4 * nf + 6 * gc) + (nf + 4 * gc) * nf
    macs_lr = 9 * (in_channels * nf + nb * 3 * rdb + nf * nf)
    macs_lr += 4 * 9 * nf * nf + 2 * 16 * 9 * nf * nf + 16 * 9 * nf * out_channels
    macs_lr += 16 * (81 * 3 * 64 + 64 * 32 + 25 * 32 * out_channels)
    return 2.0 * macs_lr / 16.0
