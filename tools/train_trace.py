"""Kernel-level timeline of one training step via torch.profiler (CUPTI): writes gpurun_out/train_trace.txt (name, us)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "climate-super-resolution_b200"))
from climsr_b200 import losses  # noqa: E402
from climsr_b200.models import ESRGANGenerator  # noqa: E402

n, h, w = 64, 64, 64
torch.manual_seed(0)
net = ESRGANGenerator(4, 1, 64, 11, 16).cuda().train()
opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-4, fused=True)
x = torch.rand(n, 4, h, w, device="cuda") * 2 - 1
e = torch.rand(n, 1, 4 * h, 4 * w, device="cuda")
m = (torch.rand(n, 1, 4 * h, 4 * w, device="cuda") > 0.3).float()
hr = torch.rand(n, 1, 4 * h, 4 * w, device="cuda") * 2 - 1


def step():
    opt.zero_grad(set_to_none=True)
    lv = losses.l1_loss(net(x, e, m), hr)
    lv.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda ev: ev.time_range.start)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "train_trace.txt"), "w") as f:
    t0 = evs[0].time_range.start
    for ev in evs:
        f.write(f"{ev.time_range.start - t0:10.1f} {ev.time_range.end - ev.time_range.start:9.1f} {ev.name[:90]}\n")
print("events", len(evs))
