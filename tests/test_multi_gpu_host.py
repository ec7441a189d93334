"""CPU tests of the N>1 paths (world_size 2, gloo): raster sharding, halo-tiled band inference and the gradient bucketer.
The arithmetic inside the bands is the CPU oracle here (test infrastructure); on the GPU the same host code drives the CUDA
generator (tests/test_gpu_parity.py::test_tiled_inference_matches_untiled)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_band_plan_covers_raster_exactly():
    from climsr_b200.tiling import band_plan, shard_indices
    for h, bands, halo in ((113, 8, 16), (113, 2, 8), (360, 8, 16), (5, 8, 2), (64, 1, 16)):
        plan = band_plan(h, bands, halo)
        assert plan[0].lo == 0 and plan[-1].hi == h
        assert all(a.hi == b.lo for a, b in zip(plan, plan[1:]))
        assert all(b.hi > b.lo for b in plan)
        assert plan[0].read_lo == 0 and plan[-1].read_hi == h           # true borders are never halo-extended
        assert all(b.lo - b.read_lo <= halo and b.read_hi - b.hi <= halo for b in plan)
    assert shard_indices(10, 1, 4) == [1, 5, 9] and shard_indices(3, 3, 4) == []
    with pytest.raises(ValueError):
        shard_indices(3, 4, 4)


def test_tiled_forward_matches_untiled_oracle():
    """Effective receptive field: halo 8 LR px reproduces the un-tiled output to fp32 noise."""
    from climsr_b200.tiling import tiled_forward_all
    from oracle import generator as og
    from oracle import synth
    sd = synth.make_state_dict(3, 1, 64, 2, 16, seed=0)
    x, elev, mask = synth.make_inputs(1, 3, 40, 24, seed=1)
    net = lambda a, b, c: og.generator_forward(sd, a, b, c)  # noqa: E731
    with torch.no_grad():
        full = net(x, elev, mask)
        errs = {}
        for halo in (0, 2, 8):
            tiled = tiled_forward_all(net, x, elev, mask, bands=3, halo=halo)
            assert tiled.shape == full.shape
            errs[halo] = float((tiled - full).abs().max())
    assert errs[8] <= 1e-5 and errs[2] < errs[0] and errs[0] > 1e-3


def test_2d_tile_grid_matches_untiled_oracle():
    """2-D halo-padded tiles (climsr_b200.tiling.tile_plan): exact cover, near-square grid choice, and the merged result equals
    the un-tiled oracle to fp32 noise at halo 8."""
    from climsr_b200.tiling import grid_shape, merge_tiles, tile_plan, tiled_forward_2d
    from oracle import generator as og
    from oracle import synth
    assert grid_shape(8, 360, 720) == (2, 4) and grid_shape(4, 360, 720) == (1, 4) or grid_shape(4, 360, 720) == (2, 2)
    assert grid_shape(1, 10, 10) == (1, 1)
    plan = tile_plan(30, 41, 2, 3, halo=8)
    assert len(plan) == 6
    cover = torch.zeros(30, 41)
    for t in plan:
        cover[t.rows.lo:t.rows.hi, t.cols.lo:t.cols.hi] += 1
        assert t.rows.read_lo <= t.rows.lo and t.cols.read_hi >= t.cols.hi
    assert bool((cover == 1).all())
    sd = synth.make_state_dict(3, 1, 64, 1, 16, seed=0)
    x, elev, mask = synth.make_inputs(1, 3, 30, 41, seed=1)
    net = lambda a, b, c: og.generator_forward(sd, a, b, c)  # noqa: E731
    with torch.no_grad():
        full = net(x, elev, mask)
        got = merge_tiles([tiled_forward_2d(net, x, elev, mask, t) for t in plan], 2, 3)
    assert got.shape == full.shape
    assert float((got - full).abs().max()) <= 1e-5


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "climate-super-resolution_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from climsr_b200.parallel import GradientBucketer
        from climsr_b200.tiling import shard_indices, tiled_forward_all
        from oracle import generator as og
        from oracle import synth
        torch.set_num_threads(2)
        # ---- (1) halo-tiled raster inference, bands sharded over ranks, result assembled on rank 0
        sd = synth.make_state_dict(3, 1, 64, 1, 16, seed=0)
        x, elev, mask = synth.make_inputs(1, 3, 24, 16, seed=1)
        net = lambda a, b, c: og.generator_forward(sd, a, b, c)  # noqa: E731

        def gather(mine):
            box = [None] * world
            dist.all_gather_object(box, mine)
            return [it for part in box for it in part]

        with torch.no_grad():
            got = tiled_forward_all(net, x, elev, mask, bands=4, halo=8, rank=rank, world=world, gather=gather)
            if rank == 0:
                full = net(x, elev, mask)
                assert float((got - full).abs().max()) <= 1e-5
            else:
                assert got is None
        # ---- (2) independent rasters: every index owned by exactly one rank
        owned = [None] * world
        dist.all_gather_object(owned, shard_indices(7, rank, world))
        assert sorted(i for part in owned for i in part) == list(range(7))
        # ---- (3) bucketed gradient all-reduce == mean of the ranks' gradients
        torch.manual_seed(0)
        lin = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.Conv2d(8, 8, 3), torch.nn.Conv2d(8, 1, 3))
        for i, p in enumerate(lin.parameters()):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        b = GradientBucketer(lin.parameters(), bucket_mb=0.001, comm_dtype=torch.float32)
        assert len(b.buckets) >= 2 and sum(len(k) for k in b.buckets) == 6
        b.allreduce()
        for i, p in enumerate(lin.parameters()):
            assert torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1)))
        # ---- (4) the overlapped path's host logic: slices of a flat gradient buffer, reduced as they "complete" (suffix first),
        # bf16 wire format averaged BEFORE the rounding, fp32 variant exact; unused parameters are refused like DDP does
        from climsr_b200.parallel import BackwardGradSync
        n = 1000
        base = torch.arange(n, dtype=torch.float32) / 7.0
        for dtype, tol in ((torch.bfloat16, 1.2e-2), (None, 1e-6)):     # two bf16 roundings of the operands + one of the sum
            flat = base * (rank + 1)
            sync = BackwardGradSync(nseg=3, comm_dtype=dtype)
            assert sync.world == 2
            pending = []
            for lo, hi in ((700, 1000), (300, 700), (0, 300)):
                sync.reduce_slice_async(flat, lo, hi, pending)
            sync.finish(flat, pending)
            want = base * 1.5
            assert float(((flat - want).abs() / (want.abs() + 1e-3)).max()) <= tol
            assert sync.last_ranges == [(700, 1000), (300, 700), (0, 300)]
        lin[0].weight.grad = None
        try:
            GradientBucketer(lin.parameters(), comm_dtype=None).allreduce()
            raise AssertionError("unused parameter was not refused")
        except RuntimeError as e:
            assert "no gradient" in str(e)
        with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
