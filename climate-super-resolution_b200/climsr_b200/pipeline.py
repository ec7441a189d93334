"""Host -> device -> host inference pipeline.

The reference's inference loop (climsr/inference/inference.py:56-82) moves one raster to the GPU, runs the model and calls
``.cpu()`` - a full device sync - per raster.  ``HostPipeline`` keeps ``depth`` batches in flight instead: the H2D copy of
batch i+1 and the D2H copy of batch i-1 run on their own streams while batch i computes, with pinned staging buffers and
events only (no host sync except when the caller consumes a result).  Throughput is then max(compute, PCIe) instead of
their sum.  The arithmetic is ``net`` (the CUDA generator); nothing here computes.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor


class _Slot:
    def __init__(self, shapes, dev):
        (xs, es, ms) = shapes
        self.x = torch.empty(xs, dtype=torch.float32, device=dev)
        self.e = torch.empty(es, dtype=torch.float32, device=dev)
        self.m = torch.empty(ms, dtype=torch.float32, device=dev)
        self.out_host = torch.empty(es, dtype=torch.float32).pin_memory()
        self.out_dev: Optional[Tensor] = None
        self.h2d_done = torch.cuda.Event()
        self.compute_done = torch.cuda.Event()
        self.d2h_done = torch.cuda.Event()
        self.busy = False


class HostPipeline:
    """``submit(x, elev, mask)`` (pinned host tensors) enqueues one batch and returns the result of the PREVIOUS batch (a
    pinned host tensor, valid until the next ``submit``) or None for the first call; ``drain()`` returns what is left."""

    def __init__(self, net, batch_shape: Tuple[int, int, int, int], device=None, depth: int = 2):
        n, c, h, w = batch_shape
        self.net = net
        self.dev = torch.device(device) if device is not None else next(net.parameters()).device
        shapes = ((n, c, h, w), (n, 1, 4 * h, 4 * w), (n, 1, 4 * h, 4 * w))
        self.slots: List[_Slot] = [_Slot(shapes, self.dev) for _ in range(depth)]
        self.h2d = torch.cuda.Stream(device=self.dev)
        self.d2h = torch.cuda.Stream(device=self.dev)
        self.i = 0

    def submit(self, x: Tensor, elev: Tensor, mask: Tensor) -> Optional[Tensor]:
        k = len(self.slots)
        s = self.slots[self.i % k]
        prev = self.slots[(self.i - 1) % k] if self.i > 0 else None
        self.i += 1
        compute = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.h2d):
            if s.busy:
                self.h2d.wait_event(s.compute_done)       # the slot's device inputs were last read by that forward
            s.x.copy_(x, non_blocking=True)
            s.e.copy_(elev, non_blocking=True)
            s.m.copy_(mask, non_blocking=True)
            s.h2d_done.record(self.h2d)
        compute.wait_event(s.h2d_done)
        with torch.no_grad():
            s.out_dev = self.net(s.x, s.e, s.m)
        s.compute_done.record(compute)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(s.compute_done)
            s.out_dev.record_stream(self.d2h)
            s.out_host.copy_(s.out_dev, non_blocking=True)
            s.d2h_done.record(self.d2h)
        s.busy = True
        # Hand back the PREVIOUS batch: its buffer belongs to another slot, so nothing enqueued above can touch it, and the
        # host waits for it while this batch's copies and kernels are already queued behind it.
        if prev is None or prev is s or not prev.busy:
            return None
        prev.d2h_done.synchronize()
        prev.busy = False
        return prev.out_host

    def drain(self) -> List[Tensor]:
        outs = []
        k = len(self.slots)
        for j in range(k):                                # oldest first
            s = self.slots[(self.i + j) % k]
            if s.busy:
                s.d2h_done.synchronize()
                outs.append(s.out_host)
                s.busy = False
        return outs


class RasterPipeline:
    """The reference's full-raster inference loop (climsr/inference/inference.py:56-82) with everything between the raw LR
    values and the denormalised, land-masked result on the GPU.

    Per batch of ``n`` rasters the reference normalises on the CPU (dataset), uploads LR + elevation + mask, runs the model,
    calls ``.cpu()`` (a device sync) and denormalises / NaN-masks with numpy, one raster at a time.  Here the static
    elevation and land mask are uploaded ONCE; ``submit(raw, mins, maxes)`` (pinned host tensors) enqueues: H2D of the raw
    rasters (4 B per LR pixel instead of 12 B per LR + 8 B per HR pixel) -> ``MinMaxScaler.normalize`` + channel concat ->
    generator -> ``MinMaxScaler.denormalize`` + NaN mask -> D2H, on three streams with ``depth`` batches in flight, and
    returns the PREVIOUS batch's result (pinned host tensor, valid until the next ``submit``)."""

    def __init__(self, net, n: int, h: int, w: int, elev: Tensor, mask: Tensor, elev_lr: Optional[Tensor] = None,
                 mask_lr: Optional[Tensor] = None, feature_range: Tuple[float, float] = (-1.0, 1.0), device=None, depth: int = 2):
        from .normalization import MinMaxScaler
        self.net = net
        self.dev = torch.device(device) if device is not None else next(net.parameters()).device
        self.n, self.h, self.w = n, h, w
        H, W = 4 * h, 4 * w
        self.scaler = MinMaxScaler(feature_range=feature_range)
        self.mask1 = mask.reshape(1, 1, H, W).to(self.dev).float().contiguous()
        self.elev_b = elev.reshape(1, 1, H, W).to(self.dev).float().expand(n, 1, H, W).contiguous()
        self.mask_b = self.mask1.expand(n, 1, H, W).contiguous()
        self.extras = [t.reshape(h, w).to(self.dev).float().contiguous() for t in (elev_lr, mask_lr) if t is not None]
        self.slots = []
        for _ in range(depth):
            s = type("Slot", (), {})()
            s.raw = torch.empty((n, h, w), dtype=torch.float32, device=self.dev)
            s.mn = torch.empty((n,), dtype=torch.float64, device=self.dev)
            s.mx = torch.empty((n,), dtype=torch.float64, device=self.dev)
            s.out_host = torch.empty((n, 1, H, W), dtype=torch.float32).pin_memory()
            s.out_dev = None
            s.h2d_done, s.compute_done, s.d2h_done = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
            s.busy = False
            self.slots.append(s)
        self.h2d = torch.cuda.Stream(device=self.dev)
        self.d2h = torch.cuda.Stream(device=self.dev)
        self.i = 0

    def submit(self, raw: Tensor, mins: Tensor, maxes: Tensor) -> Optional[Tensor]:
        k = len(self.slots)
        s = self.slots[self.i % k]
        prev = self.slots[(self.i - 1) % k] if self.i > 0 else None
        self.i += 1
        compute = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.h2d):
            if s.busy:
                self.h2d.wait_event(s.compute_done)
            s.raw.copy_(raw.reshape(self.n, self.h, self.w), non_blocking=True)
            s.mn.copy_(mins.reshape(-1).double(), non_blocking=True)
            s.mx.copy_(maxes.reshape(-1).double(), non_blocking=True)
            s.h2d_done.record(self.h2d)
        compute.wait_event(s.h2d_done)
        with torch.no_grad():
            x = self.scaler.normalize(s.raw, s.mn, s.mx, self.extras)
            sr = self.net(x, self.elev_b, self.mask_b)
            s.out_dev = self.scaler.denormalize(sr, s.mn, s.mx, self.mask1)
        s.compute_done.record(compute)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(s.compute_done)
            s.out_dev.record_stream(self.d2h)
            s.out_host.copy_(s.out_dev, non_blocking=True)
            s.d2h_done.record(self.d2h)
        s.busy = True
        if prev is None or prev is s or not prev.busy:
            return None
        prev.d2h_done.synchronize()
        prev.busy = False
        return prev.out_host

    def drain(self) -> List[Tensor]:
        outs = []
        k = len(self.slots)
        for j in range(k):
            s = self.slots[(self.i + j) % k]
            if s.busy:
                s.d2h_done.synchronize()
                outs.append(s.out_host)
                s.busy = False
        return outs
