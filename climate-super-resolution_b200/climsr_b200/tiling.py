"""Multi-GPU inference without a collective.

Two ways the generator path shards, both with NO data-path exchange step:

* many independent rasters (the reference's loop over months x variables, climsr/inference/inference.py:56-71, batch
  size 1 there): ``shard_indices`` gives rank r every world-th raster;
* one large raster (full Europe extent 113x113 LR -> 452x452 HR, or the global 360x720 grid): ``band_plan`` cuts the LR
  raster into row bands, each read with ``halo`` extra LR rows on the sides that touch another band (true image
  borders keep the convolutions' zero padding); ``tiled_forward`` runs a band on this rank and ``merge_bands`` pastes
  the cropped HR bands.  The theoretical receptive field of the trunk (167 LR px at nb=11) exceeds the Europe raster,
  so exactness by halo is impossible in principle; the measured effective field is small (halo 8 reproduces the
  un-tiled output to fp32 noise for random-init weights) - ``halo`` is a parameter and tests /
  bench.py report the error against the un-tiled run.  ``tile_plan`` / ``tiled_forward_2d`` / ``merge_tiles`` cut in both
  directions (near-square tiles carry the least halo work).

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


@dataclass(frozen=True)
class Tile2D:
    """A rectangular LR tile: rows from ``rows`` (a Band), columns from ``cols`` (a Band over the width)."""
    rows: Band
    cols: Band


def grid_shape(n_tiles: int, h: int, w: int) -> Tuple[int, int]:
    """(tiles_y, tiles_x) with tiles_y * tiles_x == n_tiles whose tiles are closest to square: least halo work per output pixel."""
    best, best_cost = (n_tiles, 1), None
    for ty in range(1, n_tiles + 1):
        if n_tiles % ty:
            continue
        tx = n_tiles // ty
        if ty > h or tx > w:
            continue
        cost = abs((h / ty) - (w / tx))
        if best_cost is None or cost < best_cost:
            best, best_cost = (ty, tx), cost
    return best


def tile_plan(h: int, w: int, tiles_y: int, tiles_x: int, halo: int = 8) -> List[Tile2D]:
    """2-D halo-padded tile grid, row-major.  A row band of 45 LR rows with a 16-row halo reads 1.7x its own rows; a
    180 x 180 tile of the global grid with the (sufficient, see module docstring) 8-px halo reads 1.19x."""
    return [Tile2D(r, c) for r in band_plan(h, tiles_y, halo) for c in band_plan(w, tiles_x, halo)]


def tiled_forward_2d(net: Callable[[Tensor, Tensor, Tensor], Tensor], x: Tensor, elev: Tensor, mask: Tensor, tile: Tile2D) -> Tensor:
    """Run one tile (halo included), crop the halo: returns HR rows [4*rows.lo, 4*rows.hi) x columns [4*cols.lo, 4*cols.hi)."""
    r, c = tile.rows, tile.cols
    xs = x[:, :, r.read_lo:r.read_hi, c.read_lo:c.read_hi].contiguous()
    es = elev[:, :, r.read_lo * SCALE:r.read_hi * SCALE, c.read_lo * SCALE:c.read_hi * SCALE].contiguous()
    ms = mask[:, :, r.read_lo * SCALE:r.read_hi * SCALE, c.read_lo * SCALE:c.read_hi * SCALE].contiguous()
    out = net(xs, es, ms)
    return out[:, :, r.crop_top:r.crop_top + r.out_rows, c.crop_top:c.crop_top + c.out_rows]


def merge_tiles(parts: Sequence[Tensor], tiles_y: int, tiles_x: int) -> Tensor:
    """Paste row-major tiles back into the raster."""
    rows = [torch.cat(list(parts[i * tiles_x:(i + 1) * tiles_x]), dim=3) for i in range(tiles_y)]
    return torch.cat(rows, dim=2)


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
