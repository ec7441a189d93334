"""Where does a training step go?  Event-timed segments (forward / loss / backward / AdamW) on cfg2-size batch."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "climate-super-resolution_b200"))
from climsr_b200 import losses  # noqa: E402
from climsr_b200.models import ESRGANGenerator  # noqa: E402

n, h, w = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 64, 64)
torch.manual_seed(0)
net = ESRGANGenerator(4, 1, 64, 11, 16).cuda().train()
opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-4, fused=True)
x = torch.rand(n, 4, h, w, device="cuda") * 2 - 1
e = torch.rand(n, 1, 4 * h, 4 * w, device="cuda")
m = (torch.rand(n, 1, 4 * h, 4 * w, device="cuda") > 0.3).float()
hr = torch.rand(n, 1, 4 * h, 4 * w, device="cuda") * 2 - 1
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
for it in range(6):
    opt.zero_grad(set_to_none=True)
    ev[0].record()
    sr = net(x, e, m)
    ev[1].record()
    lv = losses.l1_loss(sr, hr)
    ev[2].record()
    lv.backward()
    ev[3].record()
    opt.step()
    ev[4].record()
    torch.cuda.synchronize()
    gn = sum(float(p.grad.norm()) ** 2 for p in net.parameters()) ** 0.5
    print(f"it {it} loss {float(lv.detach()):.7f} |grad| {gn:.4e} fwd {ev[0].elapsed_time(ev[1]):.2f} loss {ev[1].elapsed_time(ev[2]):.2f} "
          f"bwd {ev[2].elapsed_time(ev[3]):.2f} opt {ev[3].elapsed_time(ev[4]):.2f} ms", flush=True)
