"""Training-sample assembly on device.

Reference: ``climsr/data/sr/climate_dataset.py:98-172`` builds every sample on the CPU with numpy + albumentations/cv2 -
random vertical / horizontal flip and ``np.rot90`` of the HR tile, its elevation and its land mask, the LR raster as a
nearest-neighbour resize by 1/4 (top-left pixel of every 4x4 block), ``elevation_lr`` / ``mask_lr`` the same way and
``x = cat[lr, elevation_lr, mask_lr]``.  ``make_lr_batch`` does the same for a whole batch that is already resident on the
GPU, in one pass (index work only: bit-exact against numpy/cv2, ``tests/golden/lr_input.npz``).  No CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from ._lib import CsrError, check, current_stream_ptr, lib


def aug_code(v_flip: bool, h_flip: bool, rot_k: int) -> int:
    """Augmentation code of one sample: the three draws of ``_get_training_sample`` (``climate_dataset.py:152-170``)."""
    return (1 if v_flip else 0) | (2 if h_flip else 0) | ((rot_k & 3) << 2)


def random_aug_codes(n: int, generator: Optional[torch.Generator] = None, v_flip: bool = True, h_flip: bool = True,
                     random_90_rotation: bool = True) -> Tensor:
    """Per-sample codes drawn like the reference (each transform with probability 1/2, rotation factor uniform in 0..3)."""
    r = torch.rand((n, 3), generator=generator)
    k = torch.randint(0, 4, (n,), generator=generator)
    code = (r[:, 0] > 0.5).int() * int(v_flip) + 2 * (r[:, 1] > 0.5).int() * int(h_flip)
    code = code + 4 * k.int() * (r[:, 2] > 0.5).int() * int(random_90_rotation)
    return code.int()


def make_lr_batch(hr: Tensor, elev: Tensor, mask: Tensor, codes: Optional[Tensor] = None, scale: int = 4) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """``hr, elev, mask`` (N,1,H,W) CUDA tensors, ``codes`` (N,) ints or None -> ``(x, hr', elev', mask')`` with
    ``x`` (N,3,H/scale,W/scale) the generator input and the primed tensors the augmented HR-side batch (the inputs themselves
    when ``codes`` is None)."""
    if not hr.is_cuda:
        raise CsrError("climsr_b200.data runs on CUDA (sm_100a) only; there is no CPU fallback")
    if hr.dim() != 4 or hr.shape[1] != 1 or elev.shape != hr.shape or mask.shape != hr.shape:
        raise ValueError(f"expected three (N,1,H,W) tensors, got {tuple(hr.shape)}, {tuple(elev.shape)}, {tuple(mask.shape)}")
    n, _, H, W = hr.shape
    if H % scale or W % scale:
        raise ValueError(f"H={H}, W={W} must be multiples of scale={scale}")
    a, e, m = (t.to(hr.device).contiguous().float() for t in (hr, elev, mask))
    x = torch.empty((n, 3, H // scale, W // scale), dtype=torch.float32, device=hr.device)
    if codes is None:
        outs = (None, None, None)
        cptr = None
    else:
        c = torch.as_tensor(codes, dtype=torch.int32).reshape(-1)
        if c.numel() != n:
            raise ValueError(f"expected {n} augmentation codes, got {c.numel()}")
        if H != W and bool(((c >> 2) & 1).any()):
            raise ValueError("rot90 by an odd factor needs square tiles")
        c = c.to(hr.device).contiguous()
        outs = tuple(torch.empty_like(a) for _ in range(3))
        cptr = c.data_ptr()
    with torch.cuda.device(hr.device):
        check(lib.csr_lr_input_from_hr(a.data_ptr(), e.data_ptr(), m.data_ptr(), n, H, W, scale, cptr,
                                       *(o.data_ptr() if o is not None else None for o in outs), x.data_ptr(), current_stream_ptr()),
              "csr_lr_input_from_hr")
    if codes is None:
        return x, a, e, m
    return (x,) + outs
