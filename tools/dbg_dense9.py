"""Bring-up of the nine-tap-fold dense-block kernel: one small generator forward with option 32 against the per-layer launches."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "climate-super-resolution_b200"))
from climsr_b200._lib import lib  # noqa: E402
from climsr_b200.models import ESRGANGenerator  # noqa: E402
from oracle import synth  # noqa: E402

n, in_ch, h, w, nb = [int(a) for a in sys.argv[1:6]] if len(sys.argv) > 5 else (1, 1, 9, 7, 1)
sd = synth.make_state_dict(in_ch, 1, 64, nb, 16, seed=6, gain=1.4)
x, elev, mask = synth.make_inputs(n, in_ch, h, w, seed=7)
lib.csr_set_option(31, 0)
lib.csr_set_option(11, 0)
outs = []
for fold9, dense in ((0, 0), (1, 1)):
    lib.csr_set_option(32, fold9)
    lib.csr_set_option(27, dense)
    net = ESRGANGenerator(in_ch, 1, 64, nb, 16)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    with torch.no_grad():
        a = net(x.cuda(), elev.cuda(), mask.cuda())
    torch.cuda.synchronize()
    outs.append(a.cpu())
    print("fold9", fold9, "ok", float(a.abs().max()), flush=True)
d = (outs[0] - outs[1]).abs()
print("max abs diff", float(d.max()), "scale", float(outs[0].abs().max()))
