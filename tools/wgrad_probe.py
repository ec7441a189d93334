"""Bring-up probe of the MN-major tcgen05 weight-gradient kernel: compares csr_conv2d_wgrad with torch autograd for a few shapes,
optionally sweeping the (LBO, SBO) descriptor interpretation.  Run on a B200:  python tools/wgrad_probe.py [--sweep]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "climate-super-resolution_b200"))

CASES = [
    # n, h, w, cin, cout, k, up2
    (1, 8, 14, 64, 16, 1, 0),
    (1, 8, 14, 128, 64, 1, 0),
    (1, 8, 14, 64, 16, 3, 0),
    (2, 16, 16, 80, 16, 3, 0),
    (2, 33, 45, 128, 64, 3, 0),
    (1, 20, 40, 32, 1, 5, 0),
    (1, 12, 20, 64, 64, 3, 1),
    (1, 24, 24, 192, 64, 3, 0),
]


def run_case(idx, opts):
    import torch
    import torch.nn.functional as F
    from climsr_b200 import ops
    from climsr_b200._lib import lib
    for k, v in opts.items():
        lib.csr_set_option(k, v)
    n, h, w, cin, cout, k, up2 = CASES[idx]
    g = torch.Generator().manual_seed(idx)
    x = (torch.rand((n, cin, h, w), generator=g) * 2 - 1).to(torch.bfloat16).float()
    s = 2 if up2 else 1
    gy = (torch.rand((n, cout, s * h, s * w), generator=g) * 2 - 1).to(torch.bfloat16).float()
    wt = torch.zeros((cout, cin, k, k), dtype=torch.float64, requires_grad=True)
    xin = F.interpolate(x, scale_factor=2, mode="nearest") if up2 else x
    y = F.conv2d(xin.double(), wt, None, padding=k // 2)
    (dw_ref,) = torch.autograd.grad(y, wt, gy.double())
    db_ref = gy.double().sum(dim=(0, 2, 3))
    xc = (cin + 63) // 64 * 64
    xb = torch.zeros((n, h, w, xc), dtype=torch.bfloat16)
    xb[..., :cin] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    gc = 64
    gb = torch.full((n, s * h, s * w, gc), 3.0, dtype=torch.bfloat16)      # channels >= cout must never be used
    gb[..., :cout] = gy.permute(0, 2, 3, 1).to(torch.bfloat16)
    dw, db = ops.conv2d_wgrad(xb.cuda(), gb.cuda(), (cout, cin, k, k), in_up2=bool(up2), scale=0.5)
    torch.cuda.synchronize()
    ew = float((dw.cpu().double() * 2 - dw_ref).abs().max()) / max(1.0, float(dw_ref.abs().max()))
    eb = float((db.cpu().double() * 2 - db_ref).abs().max()) / max(1.0, float(db_ref.abs().max()))
    print(f"case {idx} {CASES[idx]} opts {opts}: rel err dw {ew:.3e} db {eb:.3e} (|dw|max {float(dw_ref.abs().max()):.3g})",
          "OK" if ew < 2e-3 and eb < 2e-3 else "FAIL", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        idx = int(sys.argv[2])
        opts = {int(a.split("=")[0]): int(a.split("=")[1]) for a in sys.argv[3:]}
        run_case(idx, opts)
        sys.exit(0)
    sweeps = [{}]
    if "--sweep" in sys.argv:
        # key 20 a_lbo, 21 a_sbo, 22 b_lbo, 23 b_sbo
        sweeps += [{21: 19456, 20: 1024, 23: 16384, 22: 1024}, {20: 19456, 21: 1024, 22: 1024, 23: 1024}]
    cases = range(len(CASES))
    for sw in sweeps:
        for i in cases:
            args = [sys.executable, __file__, "--one", str(i)] + [f"{k}={v}" for k, v in sw.items()]
            try:
                r = subprocess.run(args, capture_output=True, text=True, timeout=120)
                out = (r.stdout + r.stderr).strip().splitlines()
                print(out[-1] if out else f"case {i}: no output (rc {r.returncode})", flush=True)
                if r.returncode != 0:
                    print("   rc", r.returncode, "|".join(out[-4:])[:400], flush=True)
            except subprocess.TimeoutExpired:
                print(f"case {i} opts {sw}: TIMEOUT", flush=True)
