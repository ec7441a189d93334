"""Time the cfg2 forward with whatever libclimsr_b200.so is in place; CSR_OPTS pairs are applied leniently (unknown keys are
ignored), so the same command can be run against libraries built from different commits on one GPU box."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "climate-super-resolution_b200"))
from climsr_b200._lib import lib  # noqa: E402
from climsr_b200.models import ESRGANGenerator  # noqa: E402

for kv in filter(None, os.environ.get("CSR_OPTS", "").split(",")):
    k, v = kv.split("=")
    lib.csr_set_option(int(k), int(v))
torch.manual_seed(0)
net = ESRGANGenerator(4, 1, 64, 11, 16).cuda().eval()
x = torch.rand(64, 4, 64, 64, device="cuda") * 2 - 1
e = torch.rand(64, 1, 256, 256, device="cuda")
m = (torch.rand(64, 1, 256, 256, device="cuda") > 0.3).float()
with torch.no_grad():
    for _ in range(3):
        net(x, e, m)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            net(x, e, m)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
print(f"{sys.argv[1] if len(sys.argv) > 1 else ''} {best:.3f} ms/step", flush=True)
