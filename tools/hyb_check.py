"""Hybrid tap fold (csr_set_option(35, bits)) against the fully folded kernels on one small generator forward + backward-free check."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "climate-super-resolution_b200")); sys.path.insert(0, ROOT)
from climsr_b200._lib import lib
from climsr_b200.models import ESRGANGenerator
from oracle import synth
sd = synth.make_state_dict(4, 1, 64, 2, 16, seed=16, gain=1.2)
x, elev, mask = synth.make_inputs(2, 4, 64, 64, seed=17)
outs = {}
for hyb in (0, 1, 2, 3, 7):
    lib.csr_set_option(35, hyb)
    net = ESRGANGenerator(4, 1, 64, 2, 16); net.load_state_dict(sd); net = net.cuda().eval()
    with torch.no_grad():
        outs[hyb] = net(x.cuda(), elev.cuda(), mask.cuda()).cpu()
    if hyb:
        print("hyb", hyb, "vs plain max abs", float((outs[0] - outs[hyb]).abs().max()), "scale", float(outs[0].abs().max()), flush=True)
