"""Kernel durations of csr_conv2d_wgrad for one layer shape at two batch sizes (torch.profiler / CUPTI): separates the fixed
cost of a launch (prologue + TMEM epilogue + reduce) from the per-tile cost.  python tools/wgrad_time.py"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "climate-super-resolution_b200"))
from climsr_b200 import ops  # noqa: E402


def run(n, h, w, cin, cout, k, x_c, g_c):
    x = (torch.rand((n, h, w, x_c), device="cuda") - 0.5).to(torch.bfloat16)
    g = (torch.rand((n, h, w, g_c), device="cuda") - 0.5).to(torch.bfloat16)
    for _ in range(2):
        ops.conv2d_wgrad(x, g, (cout, cin, k, k))
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            ops.conv2d_wgrad(x, g, (cout, cin, k, k))
        torch.cuda.synchronize()
    agg = {}
    for ev in prof.events():
        if ev.device_time > 0:
            a = agg.setdefault(ev.name[:60], [0, 0.0])
            a[0] += 1
            a[1] += ev.device_time
    print(f"== n{n} {h}x{w} cin{cin} cout{cout} k{k}: tiles of 128 px = {n * h * ((w + 29) // 30) * 32 // 128}")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"   {t / c:8.1f} us x {c // 3:2d} per call  {name}")


if __name__ == "__main__":
    for n in (64, 8, 1):
        run(n, 64, 64, 128, 64, 3, 128, 64)      # RDB conv5
    run(64, 64, 64, 64, 64, 3, 64, 64)           # trunk_conv-like
    run(16, 256, 256, 64, 64, 3, 64, 64)         # HRconv
