"""Discriminator alone (batch 16 of 128x128): GPU time (CUDA events) and host enqueue time (wall clock before the sync) of
forward, forward+backward(all gradients), forward+backward(input gradient only), CUDA path vs stock PyTorch bf16 autocast."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "climate-super-resolution_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from climsr_b200.models.discriminator import Discriminator  # noqa: E402
from bench_gan_step import make_discriminator  # noqa: E402


def timeit(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host = 0.0
    gpu = 0.0
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        fn()
        e1.record()
        host += time.perf_counter() - t0
        torch.cuda.synchronize()
        gpu += e0.elapsed_time(e1)
    return gpu / reps, host / reps * 1e3


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    x = (torch.rand(n, 1, 128, 128, device=dev) * 2 - 1)
    xg = x.clone().requires_grad_(True)
    for name, D, ac in (("cuda", Discriminator().to(dev).train(), False),
                        ("torch", make_discriminator().to(dev).to(memory_format=torch.channels_last).train(), True)):
        def run(inp):
            if ac:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return D(inp).float()
            return D(inp)

        def fwd():
            with torch.no_grad():
                run(x)

        def fwd_bwd_params():
            for p in D.parameters():
                p.grad = None
            run(x).sum().backward()

        def fwd_bwd_input():
            for p in D.parameters():
                p.requires_grad_(False)
            xg.grad = None
            run(xg).sum().backward()
            for p in D.parameters():
                p.requires_grad_(True)

        for label, fn in (("forward (no_grad)", fwd), ("forward + backward (parameter grads)", fwd_bwd_params),
                          ("forward + backward (input grad only)", fwd_bwd_input)):
            g, h = timeit(fn)
            print(f"{name:6s} batch {n:3d}  {label:40s} gpu {g:7.3f} ms   host enqueue {h:7.3f} ms", flush=True)


if __name__ == "__main__":
    main()
