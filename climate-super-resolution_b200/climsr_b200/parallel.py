"""Data-parallel training of the generator: gradient all-reduce over torch.distributed (NCCL on NVLink/NVSwitch in
production, gloo in the CPU tests), the DDP-equivalent of SURVEY.md section 8e.

The reference trains through Lightning's DDP plugin (conf/trainer/benchmark.yaml:4); its exchange step is one gradient
all-reduce per optimizer, started from backward hooks so that it overlaps the rest of the backward.  Here the whole
generator backward is ONE autograd node; the equivalent is ``BackwardGradSync`` (installed by ``attach_ddp``):

* the node runs its backward in ``nseg`` segments (csr_plan_backward_flat_seg; default 2: at the cfg3 step two slices hide the
  exchange completely, four cost 0.2 ms more in launches - 6.91 vs 7.11 ms on 8 GPUs).  Layers finish in reverse order, so after
  each segment a growing suffix of the plan's flat fp32 gradient buffer is final;
* that slice goes straight from the flat buffer into one persistent bf16 exchange buffer (csr_grad_pack_bf16, pre-scaled by
  1 / world so the sum over ranks stays in range; no torch.cat / cast / per-parameter copies), is all-reduced on a side
  stream behind an event (no host sync) while the next segment's kernels run, and is written back as fp32
  (csr_grad_unpack_bf16).  ``comm_dtype=None`` all-reduces the fp32 slice in place instead (the reference's numerics);
* the backward kernels are persistent grids of one CTA per SM, so a concurrent collective used to push every overlapping
  launch into a second wave (round 1: 12.1 vs 7.3 ms per cfg3 step on two GPUs).  ``attach_ddp`` therefore makes training
  plans leave ``reserve_sms`` SMs free (csr_set_option(30, k)) and the caller caps NCCL to as many CTAs
  (``NCCL_MAX_CTAS``, set before the communicator is created - bench.py does): compute and collective no longer share an SM.

``GradientBucketer`` (all-reduce after backward, any module) is kept for parameters that do not come from the generator
node - the discriminator in GAN training.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

DEFAULT_RESERVE_SMS = 4


def configure_nccl_for_overlap(reserve_sms: int = DEFAULT_RESERVE_SMS) -> None:
    """Call BEFORE init_process_group: caps NCCL's CTAs to the SMs the training plans leave free."""
    os.environ.setdefault("NCCL_MAX_CTAS", str(max(1, reserve_sms)))
    os.environ.setdefault("NCCL_MIN_CTAS", "1")


def _world(group) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class GradientBucketer:
    """Bucketed all-reduce of ``.grad`` after backward (reverse parameter order == the order gradients become final)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 4.0, comm_dtype: Optional[torch.dtype] = torch.bfloat16,
                 process_group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = process_group
        self.comm_dtype = comm_dtype
        limit = int(bucket_mb * (1 << 20))
        esize = torch.tensor([], dtype=comm_dtype or torch.float32).element_size()
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * esize
            if cur and cur_bytes + nbytes > limit:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(cur)
        self._pending = []
        self._stream = None

    @property
    def world(self) -> int:
        return _world(self.group)

    def bucket_bytes(self) -> List[int]:
        esize = torch.tensor([], dtype=self.comm_dtype or torch.float32).element_size()
        return [sum(p.numel() for p in b) * esize for b in self.buckets]

    def allreduce_async(self) -> None:
        """Start the all-reduce of every bucket (call right after loss.backward())."""
        if self.world == 1:
            return
        missing = [i for i, p in enumerate(self.params) if p.grad is None]
        if missing:
            # like DDP without find_unused_parameters: ranks must agree on the set of reduced tensors, and inventing zero
            # gradients would change the optimizer's behaviour (AdamW decay / moments of unused parameters)
            raise RuntimeError(f"GradientBucketer: {len(missing)} parameters have no gradient (first index {missing[0]}); "
                               "unused parameters are not supported")
        dev = self.params[0].device
        on_cuda = dev.type == "cuda"
        if on_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=dev)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(dev))
            self._stream.wait_event(ready)
        inv = 1.0 / self.world
        ctx = torch.cuda.stream(self._stream) if on_cuda else _null()
        with ctx:
            for bucket in self.buckets:
                # one concatenation + one cast per bucket (not per parameter); averaged BEFORE the rounding to comm_dtype
                flat = torch.cat([p.grad.reshape(-1) for p in bucket]).mul_(inv)
                if self.comm_dtype is not None and flat.dtype != self.comm_dtype:
                    flat = flat.to(self.comm_dtype)
                work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                self._pending.append((bucket, flat, work))

    def wait(self) -> None:
        """Join the collectives and write the averaged gradients back into p.grad."""
        if not self._pending:
            return
        dev = self.params[0].device
        on_cuda = dev.type == "cuda"
        ctx = torch.cuda.stream(self._stream) if on_cuda else _null()
        with ctx:
            for bucket, flat, work in self._pending:
                work.wait()
                avg = flat.to(torch.float32)
                parts = torch.split(avg, [p.numel() for p in bucket])
                torch._foreach_copy_([p.grad for p in bucket], [t.view_as(p) for t, p in zip(parts, bucket)])   # one multi-tensor kernel
        if on_cuda:
            done = torch.cuda.Event()
            done.record(self._stream)
            torch.cuda.current_stream(dev).wait_event(done)
        self._pending = []

    def allreduce(self) -> None:
        self.allreduce_async()
        self.wait()


class BackwardGradSync:
    """Gradient all-reduce overlapped with the generator's backward (see the module docstring).

    ``ESRGANGenerator.set_grad_sync(sync)`` (or ``attach_ddp``) makes the generator's autograd node call
    ``reduce_slice_async`` after every backward segment and ``finish`` at the end; the node then returns gradients that are
    already averaged over the process group, so nothing is left to do after ``loss.backward()``.
    """

    def __init__(self, nseg: int = 2, comm_dtype: Optional[torch.dtype] = torch.bfloat16, process_group=None):
        if comm_dtype not in (None, torch.bfloat16, torch.float32):
            raise ValueError("comm_dtype must be torch.bfloat16 (wire format of SURVEY 8e) or None / torch.float32 (exact fp32 sum)")
        self.nseg = max(1, int(nseg))
        self.comm_dtype = None if comm_dtype == torch.float32 else comm_dtype
        self.group = process_group
        self._stream = None
        self._comm = None                 # persistent bf16 exchange buffer, same indexing as the flat gradient buffer
        self.last_ranges: List[tuple] = []
        self.exposed_event_pairs: List[tuple] = []

    @property
    def world(self) -> int:
        return _world(self.group)

    def stream(self, dev) -> "torch.cuda.Stream":
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=dev)
        return self._stream

    def _comm_buffer(self, flat: torch.Tensor) -> torch.Tensor:
        if self._comm is None or self._comm.numel() != flat.numel() or self._comm.device != flat.device:
            self._comm = torch.empty(flat.numel(), dtype=torch.bfloat16, device=flat.device)
        return self._comm

    def reduce_slice_async(self, flat: torch.Tensor, lo: int, hi: int, pending: list) -> None:
        """Called by the autograd node right after a segment: flat[lo:hi] is final on the compute stream."""
        if hi <= lo or self.world == 1:
            return
        dev = flat.device
        inv = 1.0 / self.world
        if dev.type != "cuda":                                   # gloo tests of the host logic
            part = flat[lo:hi]
            buf = (part * inv).to(self.comm_dtype) if self.comm_dtype is not None else part.mul_(inv)
            work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            pending.append((lo, hi, buf, work))
            return
        from ._lib import check, lib
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        comm = self.stream(dev)
        with torch.cuda.stream(comm):
            comm.wait_event(ready)
            if self.comm_dtype is not None:
                buf = self._comm_buffer(flat)[lo:hi]
                check(lib.csr_grad_pack_bf16(flat.data_ptr() + 4 * lo, buf.data_ptr(), hi - lo, inv, comm.cuda_stream), "csr_grad_pack_bf16")
            else:
                buf = flat[lo:hi].mul_(inv)
            work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            pending.append((lo, hi, buf, work))

    def finish(self, flat: torch.Tensor, pending: list) -> None:
        if not pending:
            return
        dev = flat.device
        self.last_ranges = [(lo, hi) for lo, hi, _, _ in pending]
        if dev.type != "cuda":
            for lo, hi, buf, work in pending:
                work.wait()
                if buf.data_ptr() != flat[lo:hi].data_ptr():
                    flat[lo:hi].copy_(buf.to(torch.float32))
            return
        from ._lib import check, lib
        comm = self.stream(dev)
        with torch.cuda.stream(comm):
            for lo, hi, buf, work in pending:
                work.wait()
                if self.comm_dtype is not None:
                    check(lib.csr_grad_unpack_bf16(buf.data_ptr(), flat.data_ptr() + 4 * lo, hi - lo, 1.0, comm.cuda_stream), "csr_grad_unpack_bf16")
            done = torch.cuda.Event()
            done.record(comm)
        torch.cuda.current_stream(dev).wait_event(done)
        flat.record_stream(comm)


def attach_ddp(generator, nseg: int = 2, comm_dtype: Optional[torch.dtype] = torch.bfloat16, process_group=None,
               reserve_sms: int = DEFAULT_RESERVE_SMS) -> Optional[BackwardGradSync]:
    """Data-parallel training of ``generator`` (a climsr_b200 ESRGANGenerator): install the overlapped gradient all-reduce and
    make the training plans created from now on leave ``reserve_sms`` SMs to the collective.  Returns the sync object, or
    None in a single-process run (nothing to exchange).  Call ``configure_nccl_for_overlap`` before init_process_group."""
    if _world(process_group) == 1:
        return None
    params = list(generator.parameters())
    if params and params[0].is_cuda:
        from ._lib import check, lib
        check(lib.csr_set_option(30, int(reserve_sms)), "csr_set_option(30)")
    for key in list(getattr(generator, "_plans", {}).keys()):      # plans made before the reservation own all SMs: rebuild them lazily
        generator._evict(key)
    sync = BackwardGradSync(nseg=nseg, comm_dtype=comm_dtype, process_group=process_group)
    generator.set_grad_sync(sync)
    return sync


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
