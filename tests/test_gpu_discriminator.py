"""GPU parity tests of the discriminator path: climsr_b200.models.Discriminator against the golden
vectors of the UNMODIFIED reference module (tests/golden/discriminator.npz, oracle/make_golden.py) and against its own
nn.Sequential parameter containers called as stock PyTorch fp32 on the CPU (which IS the reference's arithmetic,
climsr/models/discriminator.py:42-46).

Tolerances: activations are bf16 (10 conv layers + 4 BatchNorms), accumulation fp32: scores within 3e-2 of the fp32 logits,
losses within 1e-2.  Gradients: LeakyReLU(0.01) has a derivative jump of 0.99 at zero, so the ~0.7 % of pre-activations whose
sign the bf16 forward flips (|value| below the rounding error) change the gradient by up to 25 % in relative L2 - for ANY
bf16 forward, the reference's own included.  The backward GRAPH is therefore checked against stock PyTorch run with the
LeakyReLU masks of our forward (a bf16-emulating reference: same masks, fp32 arithmetic): relative L2 <= 5e-2, cosine >=
0.998 for every parameter and for the input; the golden gradients of the reference module bound the end-to-end drift."""
import copy
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref_forward(dcpu, x):
    """The reference forward on the module's own containers (stock PyTorch)."""
    f = dcpu.feature_extraction(x)
    return dcpu.classification(f.view(f.size(0), -1))


def _make(seed=0):
    from climsr_b200.models.discriminator import Discriminator
    from oracle import synth
    d = Discriminator()
    d.load_state_dict(synth.make_discriminator_state_dict(seed=seed), strict=True)     # names / shapes / order of the reference
    return d


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12)), float(F.cosine_similarity(a.flatten(), b.flatten(), dim=0))


def test_scores_losses_and_gradients_match_reference_golden(golden_dir):
    from oracle import discriminator as od
    z = np.load(os.path.join(golden_dir, "discriminator.npz"))
    hr, sr = torch.from_numpy(z["hr"]), torch.from_numpy(z["sr"])
    d = _make().cuda().train()
    assert list(d.state_dict().keys()) == [str(k) for k in z["names"]]
    s_real, s_fake = d(hr.cuda()), d(sr.cuda())
    assert s_real.shape == (3, 1)
    assert float((s_real.detach().cpu() - torch.from_numpy(z["s_real"])).abs().max()) <= 3e-2
    assert float((s_fake.detach().cpu() - torch.from_numpy(z["s_fake"])).abs().max()) <= 3e-2
    loss_g, loss_d = od.relativistic_losses(s_real, s_fake)                 # pl_gan.py:31-39, 52-59 on the (3,1) logits
    assert abs(float(loss_g.detach()) - float(z["loss_g"])) <= 1e-2
    assert abs(float(loss_d.detach()) - float(z["loss_d"])) <= 1e-2
    loss_d.backward()
    params = dict(d.named_parameters())
    for name, key, tol, cmin in (("feature_extraction.1.weight", "g_first", 0.4, 0.92), ("classification.1.weight", "g_last", 6e-2, 0.995)):
        err, cos = _rel(params[name].grad.cpu(), torch.from_numpy(z[key]))
        assert err <= tol and cos >= cmin, (name, err, cos)       # first layer: sign-flip drift of nine LeakyReLUs, see the module docstring
    # two train-mode forwards updated the running statistics twice
    assert int(d.feature_extraction[3].num_batches_tracked) == 2
    # eval mode: running statistics of the checkpoint
    d2 = _make().cuda().eval()
    with torch.no_grad():
        s_eval = d2(hr.cuda()).cpu()
    assert float((s_eval - torch.from_numpy(z["s_eval"])).abs().max()) <= 3e-2 * max(1.0, float(np.abs(z["s_eval"]).max()))


def _logical(buf, off, step, n_log):
    return buf[:, off:off + step * n_log:step, off:off + step * n_log:step, :].float().permute(0, 3, 1, 2).cpu()


def test_all_gradients_and_running_statistics_match_stock_pytorch():
    """Every parameter gradient, the INPUT gradient (what the generator's adversarial loss back-propagates, pl_gan.py:31-47)
    and the BatchNorm running statistics, against the same module's containers run as stock PyTorch fp32 on the CPU with the
    LeakyReLU masks taken from OUR forward (see the module docstring)."""
    import torch.nn as nn
    from climsr_b200.models import discriminator as D
    d = _make(seed=3)
    ref = copy.deepcopy(d).train()
    d = d.cuda().train()
    g = torch.Generator().manual_seed(7)
    x = torch.rand((4, 1, 128, 128), generator=g) * 2 - 1
    w = torch.randn((4, 1), generator=g)
    conv_outs = []
    orig_conv = D._conv

    def spy(p, wt, b, slope, cache=None):
        o = orig_conv(p, wt, b, slope, cache)
        conv_outs.append(o)
        return o
    D._conv = spy
    try:
        xg = x.cuda().requires_grad_(True)
        out = d(xg)
        (out * w.cuda()).sum().backward()
    finally:
        D._conv = orig_conv
    assert len(conv_outs) == 10
    # logical outputs of the ten convs: stride-1 convs = interior, stride-2 convs = every second interior pixel, valid convs = interior
    sizes = [(1, 128), (2, 64), (1, 64), (2, 32), (1, 32), (2, 16), (1, 16), (2, 8), (1, 6), (1, 4)]
    masks = [_logical(o, 1, st, nl) > 0 for o, (st, nl) in zip(conv_outs, sizes)]
    xr = x.clone().requires_grad_(True)
    h, k = xr, 0
    for m in ref.feature_extraction:
        if isinstance(m, nn.LeakyReLU):
            assert masks[k].shape == h.shape
            h = h * torch.where(masks[k], torch.ones(()), torch.full((), m.negative_slope))
            k += 1
        else:
            if isinstance(m, nn.Conv2d):
                pass
            h = m(h)
    assert k == 9
    outr = ref.classification(h.view(h.size(0), -1))
    (outr * w).sum().backward()
    assert float((out.detach().cpu() - outr.detach()).abs().max()) <= 3e-2
    err, cos = _rel(xg.grad.cpu(), xr.grad)
    assert err <= 5e-2 and cos >= 0.998, ("input", err, cos)
    for (name, p), (_, q) in zip(d.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, name
        err, cos = _rel(p.grad.cpu(), q.grad)
        assert err <= 5e-2 and cos >= 0.998, (name, err, cos)       # includes the forward drift of the bf16 activations (~2-3 %)
    for (name, b), (_, c) in zip(d.named_buffers(), ref.named_buffers()):
        if name.endswith("num_batches_tracked"):
            assert int(b) == int(c)
        else:
            assert float((b.cpu() - c).abs().max()) <= 2e-2 * max(1.0, float(c.abs().max())), name


def test_frozen_discriminator_still_returns_the_input_gradient():
    """Lightning's toggle_optimizer freezes the discriminator while the generator trains: no parameter gradients, but the
    gradient w.r.t. the input must flow."""
    d = _make(seed=1).cuda().train()
    for p in d.parameters():
        p.requires_grad_(False)
    x = (torch.rand((2, 1, 128, 128)) * 2 - 1).cuda().requires_grad_(True)
    d(x).sum().backward()
    assert x.grad is not None and float(x.grad.abs().max()) > 0
    assert all(p.grad is None for p in d.parameters())
    with pytest.raises(ValueError):
        d(torch.zeros((1, 1, 64, 64), device="cuda"))


def test_gan_training_step_call_pattern():
    """The two-optimizer call pattern of GANLightningModule.training_step (pl_gan.py:63-95) through climsr_b200.task: generator
    step (optimizer_idx 0: adversarial + pixel loss through the frozen discriminator into the generator), discriminator step
    (optimizer_idx 1), validation step with the masked metrics - batch dicts with the reference's keys."""
    from climsr_b200.models import ESRGANGenerator
    from climsr_b200.task import SuperResolutionTask
    torch.manual_seed(0)
    gen = ESRGANGenerator(4, 1, 64, 1, 16).cuda()
    task = SuperResolutionTask(gen, discriminator=_make(seed=2).cuda()).train()
    g = torch.Generator().manual_seed(5)
    n = 2
    mask = (torch.rand((n, 1, 128, 128), generator=g) > 0.3).float()
    batch = {"lr": (torch.rand((n, 4, 32, 32), generator=g) * 2 - 1).cuda(), "hr": (torch.rand((n, 1, 128, 128), generator=g) * 2 - 1).cuda(),
             "elevation": (torch.rand((n, 1, 128, 128), generator=g) * mask).cuda(), "mask": mask.cuda(),
             "original_data": (torch.rand((n, 1, 128, 128), generator=g) * 40 - 10).cuda(),
             "min": torch.tensor([-30.0, -25.0], dtype=torch.float64).cuda(), "max": torch.tensor([35.0, 40.0], dtype=torch.float64).cuda()}
    opt_g = torch.optim.AdamW(task.generator.parameters(), lr=1e-4, fused=True)
    opt_d = torch.optim.AdamW(task.discriminator.parameters(), lr=1e-4, fused=True)
    w_before = task.generator.conv_last.weight.detach().clone()
    d_before = task.discriminator.classification[1].weight.detach().clone()
    # optimizer 0 (Lightning toggles requires_grad of the other optimizer's parameters off)
    for p in task.discriminator.parameters():
        p.requires_grad_(False)
    out = task.training_step(batch, 0, optimizer_idx=0)
    assert set(out["log"]) == {"train/perceptual_loss", "train/adversarial_loss", "train/pixel_level_loss", "train/loss_G"}
    opt_g.zero_grad()
    out["loss"].backward()
    assert all(p.grad is not None for p in task.generator.parameters())
    opt_g.step()
    for p in task.discriminator.parameters():
        p.requires_grad_(True)
    # optimizer 1
    for p in task.generator.parameters():
        p.requires_grad_(False)
    out = task.training_step(batch, 0, optimizer_idx=1)
    opt_d.zero_grad()
    out["loss"].backward()
    opt_d.step()
    for p in task.generator.parameters():
        p.requires_grad_(True)
    assert not torch.equal(w_before, task.generator.conv_last.weight.detach())
    assert not torch.equal(d_before, task.discriminator.classification[1].weight.detach())
    task.eval()
    val = task.validation_step(batch)
    assert "val/psnr" in val and "val/loss_G" in val and "sr" not in val
    assert all(torch.isfinite(v).all() for v in val.values())


def test_graph_replay_matches_the_eager_path_and_follows_weight_updates():
    """From the third call of a kind on, forward and backward replay CUDA graphs over leased static buffers (two slots while D(hr)
    and D(sr) of one GAN batch are both in flight).  Replays must reproduce the eager path - also after an optimizer step (weight
    packs are refreshed outside the graphs) - and hand out gradients that later replays do not overwrite."""
    d = _make(seed=5).cuda().train()
    e = copy.deepcopy(d)
    e.use_cuda_graphs = False
    opt_d = torch.optim.AdamW(d.parameters(), lr=3e-3, fused=True)
    g = torch.Generator().manual_seed(11)
    kept = []
    for it in range(5):
        xa = (torch.rand((4, 1, 128, 128), generator=g) * 2 - 1).cuda()
        xb = (torch.rand((4, 1, 128, 128), generator=g) * 2 - 1).cuda().requires_grad_(True)
        xb2 = xb.detach().clone().requires_grad_(True)
        w = torch.randn((4, 1), generator=g).cuda()
        outs = []
        # the eager twin starts every iteration from the graphed net's state: Adam turns last-bit differences of near-zero weight
        # gradients (the weight-gradient GEMM accumulates with atomics) into +-lr steps, so two independently stepped nets drift
        e.load_state_dict(d.state_dict())
        for net, xin in ((d, xb), (e, xb2)):
            net.zero_grad(set_to_none=True)
            sa, sb = net(xa), net(xin)                      # two calls in flight before the backward, as in pl_gan.py:51-61
            ((sa - sb.mean()) * w).sum().backward()
            outs.append((sa.detach().clone(), sb.detach().clone(), xin.grad.detach().clone(), [p.grad for p in net.parameters()]))
        opt_d.step()
        (sa_d, sb_d, dx_d, gr_d), (sa_e, sb_e, dx_e, gr_e) = outs
        # BatchNorm statistics and weight gradients are accumulated with atomics: two runs agree to the last bits of those sums, and a
        # last-bit difference can flip single bf16 roundings downstream - "equal" means far inside what a stale weight pack or a stale
        # static buffer would cause (the optimizer moves every weight by ~3e-3 per iteration)
        for a, b in ((sa_d, sa_e), (sb_d, sb_e)):
            assert float((a - b).abs().max()) <= 2e-3 * max(1.0, float(b.abs().max())), (it, float((a - b).abs().max()))
        assert float((dx_d - dx_e).norm()) <= 1e-2 * float(dx_e.norm()) + 1e-12, it
        for (name, _), a, b in zip(d.named_parameters(), gr_d, gr_e):
            assert float((a - b).norm()) <= 1e-2 * float(b.norm()) + 1e-9, (it, name)
        kept.append((gr_d[0], gr_d[0].clone()))
    slots = d.__dict__.get("_slots", {})
    assert slots and max(len(v) for v in slots.values()) == 2       # graphs were used, two calls in flight
    for held, copy_ in kept:                                          # gradients handed out earlier were not overwritten by later replays
        assert torch.equal(held, copy_)
    for (name, b), (_, c) in zip(d.named_buffers(), e.named_buffers()):
        assert float((b.float() - c.float()).abs().max()) <= 1e-4 * max(1.0, float(c.float().abs().max())), name
    e.load_state_dict(d.state_dict())
    with torch.no_grad():                                             # inference calls replay their own (no-save) graphs
        xs = (torch.rand((4, 1, 128, 128), generator=g) * 2 - 1).cuda()
        for _ in range(4):
            a, b = d(xs), e(xs)
            assert float((a - b).abs().max()) <= 2e-3 * max(1.0, float(b.abs().max()))
