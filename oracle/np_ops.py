"""numpy (float64) restatement of the ATen ops the reference generator calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Independent of torch's
convolution so that oracle/generator.py (graph in torch.nn.functional) can be
cross-checked: the two must agree to fp32 round-off before either is trusted.
Only meant for small cases.

Published semantics restated (PyTorch 1.10 docs; call sites esrgan.py:22-27,
33-38,90-100 and srcnn.py:9-18):
  conv2d(x,w,b,stride=1,padding=p): out[n,o,y,x] = b[o] + sum_{c,i,j} w[o,c,i,j] * xpad[n,c,y+i,x+j]
  leaky_relu(x,0.2) = x if x>=0 else 0.2*x ; relu(x) = max(x,0)
  interpolate(scale_factor=2, mode="nearest"): dst[i,j] = src[i//2, j//2]
  cat(dim=1): channel concatenation in argument order.
"""
from __future__ import annotations

from typing import Dict

import numpy as np


def conv2d(x: np.ndarray, w: np.ndarray, b: np.ndarray, pad: int) -> np.ndarray:
    n, c, h, wd = x.shape
    o, c2, kh, kw = w.shape
    assert c == c2
    xp = np.zeros((n, c, h + 2 * pad, wd + 2 * pad), dtype=np.float64)
    xp[:, :, pad:pad + h, pad:pad + wd] = x
    out = np.zeros((n, o, h, wd), dtype=np.float64)
    w64 = w.astype(np.float64)
    for i in range(kh):
        for j in range(kw):
            patch = xp[:, :, i:i + h, j:j + wd]                   # (n,c,h,w)
            out += np.einsum("nchw,oc->nohw", patch, w64[:, :, i, j])
    return out + b.astype(np.float64)[None, :, None, None]


def lrelu(x: np.ndarray, slope: float = 0.2) -> np.ndarray:
    return np.where(x >= 0, x, slope * x)


def up2(x: np.ndarray) -> np.ndarray:
    return x.repeat(2, axis=2).repeat(2, axis=3)


def _c(sd, name, x, pad):
    return conv2d(x, sd[name + ".weight"], sd[name + ".bias"], pad)


def rdb(sd: Dict[str, np.ndarray], p: str, x: np.ndarray) -> np.ndarray:
    x1 = lrelu(_c(sd, p + ".conv1", x, 1))
    x2 = lrelu(_c(sd, p + ".conv2", np.concatenate((x, x1), 1), 1))
    x3 = lrelu(_c(sd, p + ".conv3", np.concatenate((x, x1, x2), 1), 1))
    x4 = lrelu(_c(sd, p + ".conv4", np.concatenate((x, x1, x2, x3), 1), 1))
    x5 = _c(sd, p + ".conv5", np.concatenate((x, x1, x2, x3, x4), 1), 1)
    return x5 * 0.2 + x


def generator_forward(sd: Dict[str, np.ndarray], x: np.ndarray, elev: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """esrgan.py:89-102 + srcnn.py:13-18 in float64 numpy."""
    sd = {k: np.asarray(v, dtype=np.float64) for k, v in sd.items()}
    x = np.asarray(x, dtype=np.float64)
    fea = _c(sd, "conv_first", x, 1)
    t = fea
    i = 0
    while f"RRDB_trunk.{i}.RDB1.conv1.weight" in sd:
        p = f"RRDB_trunk.{i}"
        o = rdb(sd, p + ".RDB1", t)
        o = rdb(sd, p + ".RDB2", o)
        o = rdb(sd, p + ".RDB3", o)
        t = o * 0.2 + t
        i += 1
    fea = fea + _c(sd, "trunk_conv", t, 1)
    fea = lrelu(_c(sd, "upconv1", up2(fea), 1))
    if "upconv2.weight" in sd:
        fea = lrelu(_c(sd, "upconv2", up2(fea), 1))
    out = _c(sd, "conv_last", lrelu(_c(sd, "HRconv", fea, 1)), 1)
    z = np.concatenate([out, np.asarray(elev, np.float64), np.asarray(mask, np.float64)], 1)
    z = np.maximum(_c(sd, "srcnn.conv1", z, 4), 0)
    z = np.maximum(_c(sd, "srcnn.conv2", z, 0), 0)
    return _c(sd, "srcnn.conv3", z, 2)
