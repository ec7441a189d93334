import sys, os, torch
sys.path.insert(0, "/root/repo/climate-super-resolution_b200")
from climsr_b200.models import ESRGANGenerator
torch.manual_seed(0)
net = ESRGANGenerator(4, 1, 64, 11, 16).cuda().eval()
for n in (64, 32, 16, 8):
    x = torch.rand(n, 4, 64, 64, device="cuda") * 2 - 1
    e = torch.rand(n, 1, 256, 256, device="cuda")
    m = (torch.rand(n, 1, 256, 256, device="cuda") > 0.3).float()
    with torch.no_grad():
        for _ in range(3): net(x, e, m)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): net(x, e, m)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"batch {n}: {ms:.3f} ms/step, {ms/n*64:.3f} ms per 64 images, {n*65536/ms/1e3:.1f} Mpx/s")
