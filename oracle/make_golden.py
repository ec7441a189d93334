"""Generate tests/golden/*.npz by running the UNMODIFIED reference generator.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container,
where /root/reference exists:

    python oracle/make_golden.py

It imports ``climsr.models.esrgan.ESRGANGenerator`` from /root/reference
(read-only), feeds it seeded inputs, and stores small fixtures.  The GPU box
has no /root/reference; tests there read only the committed .npz files.

Fixtures
  gen_tiny_refinit.npz  - reference ctor + torch.manual_seed(0) default init (nb=1, gc=16, in=2 as in
                          tests/models/test_esrgan.py:10-12); FULL state_dict + inputs + output + L1 grads of
                          three representative weights + MSE-loss grads (all biases, six weights, all norms).
  gen_hydra_seeded.npz  - Hydra cfg (nb=11, gc=16, in=4) with oracle.synth.make_state_dict(seed=0) weights
                          loaded through load_state_dict(strict=True): inputs are re-derivable from seeds, only
                          the output (and its gain=4 "trained-like" variant) is stored.
  gen_default_seeded.npz- class default (nb=23, gc=32, in=4), same recipe, smaller raster.
  normalization.npz     - the reference's MinMaxScaler.normalize / .denormalize (+ NaN land mask) on seeded rasters.
  lr_input.npz          - numpy flips / rot90 + cv2 INTER_NEAREST resize (the arithmetic behind climate_dataset.py:152-172).
  gen_fulldepth.npz     - reference outputs at the BASELINE configs' own tile sizes and FULL depth: two cfg2 tiles (in=4, 64x64 LR,
                          nb=11, gc=16; default and "trained-like" gain), one cfg4 Europe raster (in=3, 113x113 LR) and one
                          class-default (nb=23, gc=32) 64x64 tile.  These exercise multi-window rows, two-tile windows with ragged
                          last rows and all 33 / 69 in-place concat-buffer rotations, which the 16x16 fixtures never do.
  rcan.npz              - the reference RCAN (default init) at a reduced depth (with its state_dict) and at the Hydra depth.
  discriminator.npz     - the reference Discriminator (default init, seed 0, train mode) + relativistic GAN losses.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from climsr.models.esrgan import ESRGANGenerator  # noqa: E402  (the reference itself)

from oracle import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def tiny_refinit():
    torch.manual_seed(0)
    net = ESRGANGenerator(in_channels=2, out_channels=1, nf=64, nb=1, gc=16).eval()
    x, elev, mask = synth.make_inputs(2, 2, 8, 8, seed=11)
    hr = synth.make_targets(torch.zeros(2, 1, 32, 32), seed=14)["hr"]
    sr = net(x, elev, mask)
    loss = torch.nn.functional.l1_loss(sr, hr)
    loss.backward()
    blob = {"x": x.numpy(), "elev": elev.numpy(), "mask": mask.numpy(), "hr": hr.numpy(),
            "sr": sr.detach().numpy(), "loss": loss.detach().numpy()}
    for k, v in net.state_dict().items():
        blob["sd/" + k] = v.numpy()
    for k in ("conv_first.weight", "RRDB_trunk.0.RDB2.conv3.weight", "srcnn.conv3.bias"):
        blob["grad/" + k] = dict(net.named_parameters())[k].grad.numpy()
    # MSE-loss gradients (smooth in sr, unlike L1's sign): every bias, six representative weights, and the norm of each
    # parameter gradient - the fixture that pins the backward path (tests/test_gpu_backward.py)
    net.zero_grad()
    hr2 = synth.make_targets(torch.zeros(2, 1, 32, 32), seed=21)["hr"] * 8.0
    loss2 = torch.nn.functional.mse_loss(net(x, elev, mask), hr2)
    loss2.backward()
    blob["hr_mse"] = hr2.numpy()
    blob["loss_mse"] = loss2.detach().numpy()
    full = ("conv_first.weight", "RRDB_trunk.0.RDB1.conv1.weight", "RRDB_trunk.0.RDB3.conv5.weight", "upconv2.weight",
            "conv_last.weight", "srcnn.conv1.weight")
    names, norms = [], []
    for k, p in net.named_parameters():
        names.append(k)
        norms.append(float(p.grad.norm()))
        if k.endswith(".bias") or k in full:
            blob["gradmse/" + k] = p.grad.numpy()
    blob["gradmse_names"] = np.array(names)
    blob["gradmse_norms"] = np.array(norms, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "gen_tiny_refinit.npz"), **blob)
    print("tiny_refinit", sr.shape, float(sr.std()), float(loss))


def seeded(name, in_ch, nb, gc, n, h, w):
    """Three weight scalings: default init (gain 1), "trained-like" O(0.1-0.5) outputs where the 1e-2 absolute budget of
    BASELINE.json is a tight test, and a chaotic "stress" gain where any bf16-I/O implementation (the reference under
    bf16 included) is only accurate to ~2 % of the output range."""
    blob = {}
    gains = GAINS[gc]
    for tag, gain in zip(("sr", "sr_trained", "sr_stress"), gains):
        sd = synth.make_state_dict(in_ch, 1, 64, nb, gc, seed=0, gain=gain)
        net = ESRGANGenerator(in_channels=in_ch, out_channels=1, nf=64, nb=nb, gc=gc, scale_factor=4).eval()
        net.load_state_dict(sd, strict=True)
        x, elev, mask = synth.make_inputs(n, in_ch, h, w, seed=1)
        with torch.no_grad():
            sr = net(x, elev, mask)
        blob[tag] = sr.numpy()
        print(name, tag, gain, sr.shape, "std", float(sr.std()), "absmax", float(sr.abs().max()))
    blob["meta"] = np.array([in_ch, nb, gc, n, h, w], dtype=np.int64)
    blob["gains"] = np.array(gains)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)


GAINS = {16: (1.0, 1.5, 1.8), 32: (1.0, 1.25, 1.4)}

# name -> (in_ch, nb, gc, n, h, w, weight seed, input seed, gain)
FULLDEPTH = {
    "cfg2_default": (4, 11, 16, 2, 64, 64, 0, 41, 1.0),
    "cfg2_trained": (4, 11, 16, 2, 64, 64, 0, 41, 1.5),
    "cfg4_trained": (3, 11, 16, 1, 113, 113, 5, 43, 1.5),
    "default64_trained": (4, 23, 32, 1, 64, 64, 7, 45, 1.15),   # gain where the reference itself, run in bf16, still keeps 1e-2 (0.008; at 1.25: 0.021)
}


def fulldepth():
    blob = {}
    for name, (in_ch, nb, gc, n, h, w, wseed, iseed, gain) in FULLDEPTH.items():
        sd = synth.make_state_dict(in_ch, 1, 64, nb, gc, seed=wseed, gain=gain)
        net = ESRGANGenerator(in_channels=in_ch, out_channels=1, nf=64, nb=nb, gc=gc, scale_factor=4).eval()
        net.load_state_dict(sd, strict=True)
        x, elev, mask = synth.make_inputs(n, in_ch, h, w, seed=iseed, blocky_mask=name.startswith("cfg4"))
        with torch.no_grad():
            sr = net(x, elev, mask)
        blob[name] = sr.numpy()
        blob[name + "_meta"] = np.array([in_ch, nb, gc, n, h, w, wseed, iseed], dtype=np.int64)
        blob[name + "_gain"] = np.array(gain)
        print("fulldepth", name, sr.shape, "std", float(sr.std()), "absmax", float(sr.abs().max()))
    np.savez_compressed(os.path.join(OUT, "gen_fulldepth.npz"), **blob)


def normalization():
    """tests/golden/normalization.npz: the reference's MinMaxScaler (climsr/data/normalization.py, imported unmodified) on
    seeded rasters with NaNs: normalize(arr, min, max) and denormalize(arr, min, max) + the NaN land mask of inference.py:75."""
    from climsr.data.normalization import MinMaxScaler
    rng = np.random.default_rng(12)
    n, h, w = 3, 19, 23
    raw = rng.uniform(-40.0, 45.0, size=(n, h, w)).astype(np.float32)
    raw[rng.uniform(size=raw.shape) < 0.2] = np.nan                        # sea pixels of the LR raster
    mins = np.array([-52.25, -38.5, -61.125], dtype=np.float64)            # pandas float64 columns in the reference
    maxes = np.array([41.5, 47.75, 36.0], dtype=np.float64)
    sc = MinMaxScaler(feature_range=(-1.0, 1.0))
    norm = np.stack([sc.normalize(raw[i], mins[i], maxes[i]) for i in range(n)])
    sr = rng.uniform(-1.1, 1.1, size=(n, 1, 4 * h, 4 * w)).astype(np.float32)
    mask = rng.uniform(size=(1, 1, 4 * h, 4 * w)) > 0.3
    post = np.empty_like(sr)
    for i in range(n):
        arr = sc.denormalize(sr[i, 0], mins[i], maxes[i])
        arr[~mask[0, 0]] = np.nan                                          # inference.py:75
        post[i, 0] = arr.astype(np.float32)
    sc01 = MinMaxScaler()                                                  # default feature_range (0, 1)
    norm01 = sc01.normalize(raw[0], mins[0], maxes[0])
    np.savez_compressed(os.path.join(OUT, "normalization.npz"), raw=raw, mins=mins, maxes=maxes, norm=norm, sr=sr,
                        mask=mask.astype(np.uint8), post=post, norm01=norm01, numpy_version=np.array(np.__version__))
    print("normalization.npz written")


def discriminator():
    """tests/golden/discriminator.npz: the reference Discriminator (climsr/models/discriminator.py, imported unmodified,
    weights = oracle.synth.make_discriminator_state_dict(0) through load_state_dict(strict=True), .train() mode) on seeded
    128x128 inputs: scores for a "real" and a "fake" batch,
    the relativistic losses of pl_gan.py:28-61 computed with torch's BCEWithLogitsLoss, and the gradient of the
    discriminator loss w.r.t. the first and last weights, plus the eval-mode scores (running statistics)."""
    from climsr.models.discriminator import Discriminator
    D = Discriminator()
    D.load_state_dict(synth.make_discriminator_state_dict(seed=0), strict=True)    # names / shapes / order must match exactly
    D.train()
    g = torch.Generator().manual_seed(31)
    hr = torch.rand((3, 1, 128, 128), generator=g) * 2 - 1
    sr = (hr + 0.1 * torch.randn((3, 1, 128, 128), generator=g)).clamp(-1, 1)
    s_real, s_fake = D(hr), D(sr)
    bce = torch.nn.BCEWithLogitsLoss()
    ones, zeros = torch.ones((3, 1)), torch.zeros((3, 1))
    rf, fr = s_real - s_fake.mean(), s_fake - s_real.mean()
    loss_g = (bce(fr, ones) + bce(rf, zeros)) / 2                       # pl_gan.py:31-39
    loss_d = (bce(rf, ones) + bce(fr, zeros)) / 2                       # pl_gan.py:52-59
    params = dict(D.named_parameters())
    first, last = "feature_extraction.1.weight", "classification.1.weight"
    g_first, g_last = torch.autograd.grad(loss_d, [params[first], params[last]])
    D.load_state_dict(synth.make_discriminator_state_dict(seed=0), strict=True)    # undo the running-statistics updates of the two forwards
    D.eval()
    with torch.no_grad():
        s_eval = D(hr)
    np.savez_compressed(os.path.join(OUT, "discriminator.npz"), hr=hr.numpy(), sr=sr.numpy(), s_real=s_real.detach().numpy(),
                        s_fake=s_fake.detach().numpy(), loss_g=float(loss_g.detach()), loss_d=float(loss_d.detach()),
                        g_first=g_first.numpy(), g_last=g_last.numpy(), s_eval=s_eval.numpy(),
                        names=np.array(list(D.state_dict().keys())))
    print("discriminator.npz written")


def rcan():
    """tests/golden/rcan.npz: the reference RCAN (climsr/models/rcan.py, imported unmodified; default init under
    torch.manual_seed, conv weights of the body scaled so that the output is O(0.3)) on seeded inputs: a reduced depth
    (2 groups x 3 blocks) with the names and per-tensor checksums of its state_dict, and the Hydra depth (10 x 20, conf/generator/rcan.yaml) whose weights are
    re-derivable from the seed (only the output is stored)."""
    from climsr.models.rcan import RCAN
    blob = {}
    for tag, (ng, nbk, n, h, w, seed) in {"small": (2, 3, 2, 20, 24, 0), "hydra": (10, 20, 1, 16, 16, 1)}.items():
        torch.manual_seed(seed)
        net = RCAN(n_resgroups=ng, n_resblocks=nbk, n_feats=64, reduction=16, scaling_factor=4, in_channels=3, out_channels=1).eval()
        x, elev, mask = synth.make_inputs(n, 3, h, w, seed=50 + seed)
        with torch.no_grad():
            sr = net(x, elev, mask)
        blob[tag] = sr.numpy()
        blob[tag + "_meta"] = np.array([ng, nbk, n, h, w, seed], dtype=np.int64)
        if tag == "small":
            # the weights are re-derivable (torch.manual_seed + the same parameter creation order): pin names and per-tensor checksums
            blob["names"] = np.array(list(net.state_dict().keys()))
            blob["sd_sum"] = np.array([float(v.double().sum()) for v in net.state_dict().values()])
            blob["sd_abs"] = np.array([float(v.double().abs().sum()) for v in net.state_dict().values()])
        print("rcan", tag, sr.shape, "std", float(sr.std()), "absmax", float(sr.abs().max()))
    np.savez_compressed(os.path.join(OUT, "rcan.npz"), **blob)


def rcan_grad():
    """tests/golden/rcan_grad.npz (groundwork for RCAN training): autograd of the UNMODIFIED reference RCAN at the
    reduced depth of rcan.npz, loss = sum(sr * w) with a seeded w.  Stored: sum and abs-sum of EVERY parameter gradient (state_dict
    order), the full gradients of four representative tensors (first conv, a channel-attention 1x1, the last upsampler conv, srcnn.conv3)
    and the gradient w.r.t. the LR input."""
    from climsr.models.rcan import RCAN
    ng, nbk, n, h, w, seed = 2, 3, 2, 20, 24, 0
    torch.manual_seed(seed)
    net = RCAN(n_resgroups=ng, n_resblocks=nbk, n_feats=64, reduction=16, scaling_factor=4, in_channels=3, out_channels=1).train()
    x, elev, mask = synth.make_inputs(n, 3, h, w, seed=50 + seed)
    x.requires_grad_(True)
    wsum = torch.randn((n, 1, 4 * h, 4 * w), generator=torch.Generator().manual_seed(77))
    sr = net(x, elev, mask)
    (sr * wsum).sum().backward()
    names = [k for k, _ in net.named_parameters()]
    grads = [p.grad for _, p in net.named_parameters()]
    blob = {"meta": np.array([ng, nbk, n, h, w, seed], dtype=np.int64), "names": np.array(names),
            "g_sum": np.array([float(g.double().sum()) for g in grads]), "g_abs": np.array([float(g.double().abs().sum()) for g in grads]),
            "dx": x.grad.numpy()}
    for k in ("head.0.weight", "body.0.body.1.body.3.conv_du.0.weight", "tail.0.2.weight", "srcnn.conv3.weight"):
        blob["g:" + k] = dict(net.named_parameters())[k].grad.numpy()
    print("rcan_grad", len(names), "tensors; |dx| max", float(x.grad.abs().max()))
    np.savez_compressed(os.path.join(OUT, "rcan_grad.npz"), **blob)


def lr_input():
    """tests/golden/lr_input.npz: numpy flips / rot90 (climate_dataset.py:152-170) and cv2.resize INTER_NEAREST - what
    albumentations' A.Resize calls (climate_dataset.py:84-92,172) - on seeded square tiles, all 16 augmentation codes."""
    import cv2
    rng = np.random.default_rng(21)
    S, s = 24, 4
    codes = np.arange(16, dtype=np.int32)
    n = len(codes)
    hr = rng.uniform(-1, 1, size=(n, 1, S, S)).astype(np.float32)
    elev = rng.uniform(-1, 1, size=(n, 1, S, S)).astype(np.float32)
    mask = (rng.uniform(size=(n, 1, S, S)) > 0.3).astype(np.float32)
    xs, hrs, els, mks = [], [], [], []
    for i, c in enumerate(codes):
        a, e, m = hr[i, 0], elev[i, 0], mask[i, 0]
        if c & 1:
            a, e, m = np.flipud(a), np.flipud(e), np.flipud(m)
        if c & 2:
            a, e, m = np.fliplr(a), np.fliplr(e), np.fliplr(m)
        k = (c >> 2) & 3
        if k:
            a, e, m = np.rot90(a, k), np.rot90(e, k), np.rot90(m, k)
        a, e, m = (np.ascontiguousarray(t) for t in (a, e, m))
        rs = lambda t: cv2.resize(t, (S // s, S // s), interpolation=cv2.INTER_NEAREST)  # noqa: E731
        xs.append(np.stack([rs(a), rs(e), rs(m.astype(np.float32))]))
        hrs.append(a); els.append(e); mks.append(m)
    # a non-square, un-augmented case (Europe-extent validation rasters are not square)
    r = rng.uniform(-1, 1, size=(20, 36)).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "lr_input.npz"), hr=hr, elev=elev, mask=mask, codes=codes, x=np.stack(xs),
                        hr_aug=np.stack(hrs)[:, None], elev_aug=np.stack(els)[:, None], mask_aug=np.stack(mks)[:, None],
                        rect=r, rect_lr=cv2.resize(r, (9, 5), interpolation=cv2.INTER_NEAREST), cv2_version=np.array(cv2.__version__))
    print("lr_input.npz written")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if "--normalization-only" in sys.argv:
        normalization()
        sys.exit(0)
    if "--lr-input-only" in sys.argv:
        lr_input()
        sys.exit(0)
    if "--discriminator-only" in sys.argv:
        discriminator()
        sys.exit(0)
    if "--rcan-only" in sys.argv:
        rcan()
        sys.exit(0)
    if "--rcan-grad-only" in sys.argv:
        rcan_grad()
        sys.exit(0)
    if "--fulldepth-only" in sys.argv:
        fulldepth()
        sys.exit(0)
    tiny_refinit()
    normalization()
    lr_input()
    discriminator()
    rcan()
    if "--tiny-only" not in sys.argv:
        seeded("gen_hydra_seeded", 4, 11, 16, 2, 16, 16)
        seeded("gen_default_seeded", 4, 23, 32, 1, 12, 12)
        fulldepth()
