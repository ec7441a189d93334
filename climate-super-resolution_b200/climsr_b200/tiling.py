"""Multi-GPU inference without a collective.

Two ways the generator path shards, both with NO data-path exchange step:

* many independent rasters (the reference's loop over months x variables, climsr/inference/inference.py:56-71, batch
  size 1 there): ``shard_indices`` gives rank r every world-th raster;
* one large raster (full Europe extent 113x113 LR -> 452x452 HR, or the global 360x720 grid): ``band_plan`` cuts the LR
  raster into row bands, each read with ``halo`` extra LR rows on the sides that touch another band (true image
  borders keep the convolutions' zero padding); ``tiled_forward`` runs a band on this rank and ``merge_bands`` pastes
  the cropped HR bands.  The theoretical receptive field of the trunk (167 LR px at nb=11) exceeds the Europe raster,
  so exactness by halo is impossible in principle; the measured effective field is small (halo 8 reproduces the
  un-tiled output to fp32 noise for random-init weights) - ``halo`` is a parameter (default 16)
  and tests report the error against the un-tiled oracle.

Pure host logic (index arithmetic + calls into ``net``); the arithmetic runs in whatever ``net`` is - the CUDA
generator in production, the CPU oracle in the CPU tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

SCALE = 4


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment of independent rasters / tile batches to ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_items, world))


@dataclass(frozen=True)
class Band:
    """LR rows [lo, hi) are this band's outputs; rows [read_lo, read_hi) are what it reads (halo included)."""
    lo: int
    hi: int
    read_lo: int
    read_hi: int

    @property
    def crop_top(self) -> int:          # HR rows to drop from the top of the band's output
        return (self.lo - self.read_lo) * SCALE

    @property
    def out_rows(self) -> int:
        return (self.hi - self.lo) * SCALE


def band_plan(h: int, bands: int, halo: int = 16) -> List[Band]:
    """Split h LR rows into `bands` contiguous bands of near-equal height (no empty band; bands <= h)."""
    if h < 1 or bands < 1 or halo < 0:
        raise ValueError("h, bands must be positive and halo non-negative")
    bands = min(bands, h)
    base, extra = divmod(h, bands)
    out, lo = [], 0
    for b in range(bands):
        hi = lo + base + (1 if b < extra else 0)
        out.append(Band(lo, hi, max(0, lo - halo), min(h, hi + halo)))
        lo = hi
    return out


def tiled_forward(net: Callable[[Tensor, Tensor, Tensor], Tensor], x: Tensor, elev: Tensor, mask: Tensor, band: Band) -> Tensor:
    """Run one band: slice LR rows [read_lo, read_hi) (and the matching HR rows of elev/mask), run the generator, crop the
    halo.  Returns the HR rows [4*lo, 4*hi)."""
    xs = x[:, :, band.read_lo:band.read_hi, :].contiguous()
    es = elev[:, :, band.read_lo * SCALE:band.read_hi * SCALE, :].contiguous()
    ms = mask[:, :, band.read_lo * SCALE:band.read_hi * SCALE, :].contiguous()
    out = net(xs, es, ms)
    return out[:, :, band.crop_top:band.crop_top + band.out_rows, :]


def merge_bands(parts: Sequence[Tensor]) -> Tensor:
    return torch.cat(list(parts), dim=2)


def tiled_forward_all(net, x: Tensor, elev: Tensor, mask: Tensor, bands: int, halo: int = 16, rank: int = 0, world: int = 1,
                      gather: Optional[Callable[[List[Tuple[int, Tensor]]], List[Tuple[int, Tensor]]]] = None) -> Optional[Tensor]:
    """Whole-raster inference over `bands` row bands.  Rank r computes bands r, r+world, ...; with world == 1 the merged
    raster is returned directly, otherwise `gather` (e.g. a torch.distributed.gather_object wrapper - result assembly, not
    part of the data path) collects (band index, tensor) pairs and rank 0 returns the merged raster."""
    plan = band_plan(x.shape[2], bands, halo)
    mine = [(i, tiled_forward(net, x, elev, mask, plan[i])) for i in shard_indices(len(plan), rank, world)]
    if world == 1:
        return merge_bands([t for _, t in mine])
    if gather is None:
        raise ValueError("world > 1 needs a gather callable")
    allparts = gather(mine)
    if rank != 0:
        return None
    allparts = sorted(allparts, key=lambda it: it[0])
    return merge_bands([t for _, t in allparts])
