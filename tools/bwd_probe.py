"""Per-parameter gradient parity of the generator training step vs the CPU oracle (fp32 autograd).  Run on a B200."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "climate-super-resolution_b200"))
from climsr_b200.models import ESRGANGenerator  # noqa: E402
from oracle import generator as og  # noqa: E402
from oracle import synth  # noqa: E402


def main(in_ch=2, nb=1, gc=16, n=2, h=16, w=16, gain=1.0):
    sd = synth.make_state_dict(in_ch, 1, 64, nb, gc, seed=0, gain=gain)
    x, elev, mask = synth.make_inputs(n, in_ch, h, w, seed=1)
    g = torch.Generator().manual_seed(7)
    hr = torch.rand((n, 1, 4 * h, 4 * w), generator=g) * 2 - 1
    sr_ref, loss_ref, grads_ref = og.generator_forward_backward(sd, x, elev, mask, hr, loss="mse")
    net = ESRGANGenerator(in_ch, 1, 64, nb, gc)
    net.load_state_dict(sd)
    net = net.cuda().train()
    sr = net(x.cuda(), elev.cuda(), mask.cuda())
    loss = torch.nn.functional.mse_loss(sr, hr.cuda())
    loss.backward()
    torch.cuda.synchronize()
    print(f"loss ref {float(loss_ref):.6f} ours {float(loss):.6f}; sr max err {float((sr.detach().cpu() - sr_ref).abs().max()):.3e}")
    worst = 0.0
    for name, p in net.named_parameters():
        ref = grads_ref[name]
        got = p.grad.detach().cpu()
        rel = float((got - ref).norm() / (ref.norm() + 1e-30))
        cos = float(torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0))
        worst = max(worst, rel)
        flag = "" if rel < 3e-2 else "   <<<<"
        print(f"{name:36s} |ref| {float(ref.norm()):.3e} rel {rel:.3e} cos {cos:.5f}{flag}")
    print("worst rel", worst)


if __name__ == "__main__":
    args = [int(a) for a in sys.argv[1:]]
    main(*args)
