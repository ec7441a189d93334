"""Secondary BASELINE.json configurations on one or more B200s (results -> gpurun_out/configs_r01.json):
  cfg1  batch 16 of 128x128 HR tiles (the reference's CPU-runnable case), Hydra generator, in=4
  cfg2d class-default RRDBNet (nb=23, gc=32), 64 tiles of 256x256 HR
  cfg4  one Europe-extent raster 113x113 -> 452x452 (in=3); and the global 360x720 -> 1440x2880 grid as halo-padded row bands
  cfg5  48 Europe-extent rasters (4 variables x 12 months) + masked loss and 16 metrics on device per raster batch
Run: python tools/bench_configs.py            (torchrun for N > 1: bands / rasters are sharded over ranks, no collective)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "climate-super-resolution_b200"))
sys.path.insert(0, ROOT)
from climsr_b200.metrics import masked_val_metrics_raw  # noqa: E402
from climsr_b200.models import ESRGANGenerator  # noqa: E402
from climsr_b200.tiling import band_plan, shard_indices, tiled_forward  # noqa: E402


def timed(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def inputs(n, c, h, w, dev):
    g = torch.Generator().manual_seed(1)
    x = (torch.rand((n, c, h, w), generator=g) * 2 - 1).to(dev)
    mask = (torch.rand((n, 1, 4 * h, 4 * w), generator=g) > 0.3).float().to(dev)
    elev = ((torch.rand((n, 1, 4 * h, 4 * w), generator=g) * 2 - 1)).to(dev) * mask
    return x, elev, mask


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    out = {"n_gpus": world}
    torch.manual_seed(0)
    with torch.no_grad():
        net4 = ESRGANGenerator(4, 1, 64, 11, 16).to(dev).eval()
        net3 = ESRGANGenerator(3, 1, 64, 11, 16).to(dev).eval()
        if world == 1:
            x, e, m = inputs(16, 4, 32, 32, dev)
            ms = timed(lambda: net4(x, e, m), 30)
            out["cfg1"] = {"ms": ms, "hr_mpx_s": 16 * 128 * 128 / ms / 1e3, "shape": "16 x (4,32,32) -> (1,128,128)"}
            netd = ESRGANGenerator(4, 1, 64, 23, 32).to(dev).eval()
            x, e, m = inputs(64, 4, 64, 64, dev)
            ms = timed(lambda: netd(x, e, m), 5)
            fl = 2275424.0
            out["cfg2_class_default"] = {"ms": ms, "hr_mpx_s": 64 * 256 * 256 / ms / 1e3, "tflops_algorithmic": 64 * 256 * 256 * fl / ms / 1e9,
                                         "shape": "nb=23 gc=32, 64 x (4,64,64)"}
            del netd
            x, e, m = inputs(1, 3, 113, 113, dev)
            ms = timed(lambda: net3(x, e, m), 30)
            out["cfg4_europe"] = {"ms": ms, "hr_mpx_s": 452 * 452 / ms / 1e3, "shape": "1 x (3,113,113) -> (1,452,452), un-tiled"}
            x, e, m = inputs(48, 3, 113, 113, dev)
            hr = (torch.rand_like(m) * 2 - 1)
            orig = hr * 30 + 5
            mn = torch.full((48,), -40.0, device=dev)
            mx = torch.full((48,), 35.0, device=dev)

            def cfg5():
                sr = net3(x, e, m)
                return masked_val_metrics_raw(sr, hr, orig, m, mn, mx)
            ms = timed(cfg5, 5)
            ms_metrics = timed(lambda: masked_val_metrics_raw(hr, hr, orig, m, mn, mx), 10)
            px = 48 * 452 * 452
            out["cfg5_48_rasters_with_metrics"] = {"ms": ms, "rasters_per_s": 48 / ms * 1e3, "hr_mpx_s": px / ms / 1e3, "metrics_ms": ms_metrics,
                                                   "metrics_gb_s": px * 28 / ms_metrics / 1e6,
                                                   "metrics_bytes_per_px": "16 (sr,hr,original,mask fp32, pass 1) + 12 (sr,hr,mask, SSIM pass)"}
            # the reference's inference loop over monthly rasters, host to host: raw LR values in, denormalised land-masked
            # rasters out (RasterPipeline: static elevation / mask resident, normalise + denormalise on device, copies overlapped)
            from climsr_b200.pipeline import HostPipeline, RasterPipeline
            import time
            nb_, reps = 48, 6
            raw = (torch.rand((nb_, 113, 113)) * 70 - 35).pin_memory()
            mins = torch.full((nb_,), -40.0, dtype=torch.float64).pin_memory()
            maxes = torch.full((nb_,), 35.0, dtype=torch.float64).pin_memory()
            rp = RasterPipeline(net3, nb_, 113, 113, e[:1].cpu(), m[:1].cpu(), x[0, 1].cpu(), x[0, 2].cpu())
            for _ in range(2):
                rp.submit(raw, mins, maxes)
            rp.drain()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                rp.submit(raw, mins, maxes)
            rp.drain()
            ms_rp = (time.perf_counter() - t0) / reps * 1e3
            hp = HostPipeline(net3, (nb_, 3, 113, 113))
            xh, eh, mh = x.cpu().pin_memory(), e.cpu().pin_memory(), m.cpu().pin_memory()
            for _ in range(2):
                hp.submit(xh, eh, mh)
            hp.drain()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                hp.submit(xh, eh, mh)
            hp.drain()
            ms_hp = (time.perf_counter() - t0) / reps * 1e3
            out["cfg5_raster_pipeline_host_to_host"] = {
                "ms_per_48_rasters": ms_rp, "rasters_per_s": nb_ / ms_rp * 1e3, "hr_mpx_s": px / ms_rp / 1e3,
                "h2d_bytes": nb_ * 113 * 113 * 4 + 2 * nb_ * 8, "d2h_bytes": px * 4,
                "same_through_HostPipeline_ms": ms_hp, "HostPipeline_h2d_bytes": nb_ * 3 * 113 * 113 * 4 + 2 * px * 4,
                "note": "wall clock over 6 batches, 2 in flight; HostPipeline re-uploads elevation + mask per batch like the reference and "
                        "returns normalised values (denormalise + NaN mask would follow on the host)"}
        # global grid as row bands: rank r takes bands r, r+world, ...
        x, e, m = inputs(1, 3, 360, 720, dev)
        bands = max(8, world)
        plan = band_plan(360, bands, 16)
        mine = shard_indices(bands, rank, world)

        def run_bands():
            return [tiled_forward(net3, x, e, m, plan[i]) for i in mine]
        ms = timed(run_bands, 5)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
        out["cfg4_global_bands"] = {"ms": ms, "hr_mpx_s": 1440 * 2880 / ms / 1e3, "bands": bands, "halo_lr_px": 16,
                                    "shape": "1 x (3,360,720) -> (1,1440,2880), bands sharded over ranks, no collective"}
        if world == 1:
            full = net3(x, e, m)
            tiled = torch.cat(run_bands(), dim=2)
            out["cfg4_global_bands"]["max_abs_vs_untiled"] = float((full - tiled).abs().max())
            ms_full = timed(lambda: net3(x, e, m), 5)
            out["cfg4_global_untiled"] = {"ms": ms_full, "hr_mpx_s": 1440 * 2880 / ms_full / 1e3}
    if rank == 0:
        print(json.dumps(out))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"configs_r01_n{world}.json"), "w") as f:
            json.dump(out, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
