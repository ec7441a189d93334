"""CPU fp32 restatement of the reference RCAN generator (climsr/models/rcan.py:50-186 + climsr/models/srcnn.py:13-18).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain functional torch on a reference ``state_dict``; pinned against the
UNMODIFIED reference module by oracle/make_golden.py (tests/golden/rcan.npz)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def rcan_forward(sd: Dict[str, Tensor], x: Tensor, elev: Tensor, mask: Tensor, n_resgroups: int, n_resblocks: int) -> Tensor:
    c = lambda name, t, pad: F.conv2d(t, sd[name + ".weight"], sd[name + ".bias"], padding=pad)  # noqa: E731
    head = c("head.0", x, 1)
    cur = head
    for g in range(n_resgroups):
        g_in = cur
        for b in range(n_resblocks):
            pre = f"body.{g}.body.{b}.body"
            r = c(pre + ".2", F.relu(c(pre + ".0", cur, 1)), 1)
            y = r.mean(dim=(2, 3), keepdim=True)                                   # CALayer (rcan.py:64-68)
            y = torch.sigmoid(c(pre + ".3.conv_du.2", F.relu(c(pre + ".3.conv_du.0", y, 0)), 0))
            cur = r * y + cur                                                      # RCAB skip (rcan.py:98-101)
        cur = c(f"body.{g}.body.{n_resblocks}", cur, 1) + g_in                     # ResidualGroup (rcan.py:131-134)
    res = c(f"body.{n_resgroups}", cur, 1) + head
    t = F.pixel_shuffle(c("tail.0.0", res, 1), 2)
    t = F.pixel_shuffle(c("tail.0.2", t, 1), 2)
    y = c("tail.1", t, 1)
    s = torch.cat([y, elev, mask], 1)
    s = F.relu(c("srcnn.conv1", s, 4))
    s = F.relu(c("srcnn.conv2", s, 0))
    return c("srcnn.conv3", s, 2)
