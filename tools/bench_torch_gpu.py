""""The kernel to beat" (BASELINE.md section 4): the reference generator graph in stock PyTorch eager + cuDNN on the
same B200, in four variants: fp32, autocast(bf16), channels_last + cudnn.benchmark + autocast(bf16), pure bf16.

Measurement tool only (imports oracle/ for the reference graph; never used by the product path):
    python tools/bench_torch_gpu.py [--workload cfg2] [--steps 10]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, flops_per_hr_pixel  # noqa: E402
from oracle import generator as og  # noqa: E402
from oracle import synth  # noqa: E402


def timed(fn, steps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    in_ch, nb, gc, tiles, h, w = WORKLOADS[a.workload]
    dev = torch.device("cuda", 0)
    sd = {k: v.to(dev) for k, v in synth.make_state_dict(in_ch, 1, 64, nb, gc, seed=0).items()}
    x, elev, mask = (t.to(dev) for t in synth.make_inputs(tiles, in_ch, h, w, seed=1))
    px = tiles * 16 * h * w
    fl = flops_per_hr_pixel(in_ch, 64, nb, gc)
    res = {}
    with torch.no_grad():
        torch.backends.cudnn.benchmark = False
        res["eager_fp32"] = timed(lambda: og.generator_forward(sd, x, elev, mask), a.steps)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            res["eager_autocast_bf16"] = timed(lambda: og.generator_forward(sd, x, elev, mask), a.steps)
        torch.backends.cudnn.benchmark = True
        sdc = {k: (v.to(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}
        xc, ec, mc = (t.contiguous(memory_format=torch.channels_last) for t in (x, elev, mask))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            res["channels_last_cudnn_benchmark_autocast_bf16"] = timed(lambda: og.generator_forward(sdc, xc, ec, mc), a.steps)
        sdb = {k: v.to(torch.bfloat16) for k, v in sdc.items()}
        xb, eb, mb = (t.to(torch.bfloat16) for t in (xc, ec, mc))
        res["channels_last_pure_bf16"] = timed(lambda: og.generator_forward(sdb, xb, eb, mb), a.steps)
    out = {"workload": a.workload, "hr_pixels_per_step": px,
           "variants": {k: {"ms_per_step": v, "mpixel_per_s": px / v / 1e3, "tflops_algorithmic": px * fl / v / 1e9}
                        for k, v in res.items()}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
