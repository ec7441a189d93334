"""Per-kernel time of the masked-metrics step at cfg5 size (48 x 452 x 452) and at cfg2 size (64 x 256 x 256)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "climate-super-resolution_b200"))
from climsr_b200.metrics import masked_val_metrics_raw  # noqa: E402

for n, H, W in ((48, 452, 452), (64, 256, 256)):
    sr = torch.rand(n, 1, H, W, device="cuda") * 2 - 1
    hr = sr + 0.05 * torch.randn_like(sr)
    m = (torch.rand_like(sr) > 0.3).float()
    orig = hr * 30 + 5
    mn = torch.full((n,), -40.0, device="cuda")
    mx = torch.full((n,), 35.0, device="cuda")
    for _ in range(3):
        masked_val_metrics_raw(sr, hr, orig, m, mn, mx)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        masked_val_metrics_raw(sr, hr, orig, m, mn, mx)
        torch.cuda.synchronize()
    px = n * H * W
    print(f"== {n} x {H} x {W}: {px/1e6:.1f} Mpx")
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and "csr" in ev.name:
            us = ev.time_range.end - ev.time_range.start
            print(f"  {ev.name[:50]:50s} {us:8.1f} us")
