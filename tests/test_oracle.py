"""Pin the oracle: golden vectors from the reference module, numpy cross-check, metric KATs."""
import os

import numpy as np
import pytest
import torch

from oracle import generator as og
from oracle import metrics as om
from oracle import np_ops, synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_tiny_refinit_matches_reference_forward_and_grads(golden_dir):
    z = _load(golden_dir, "gen_tiny_refinit.npz")
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    x, elev, mask, hr = (torch.from_numpy(z[k]) for k in ("x", "elev", "mask", "hr"))
    sr, loss, grads = og.generator_forward_backward(sd, x, elev, mask, hr, "l1")
    assert sr.shape == (2, 1, 32, 32)
    assert np.abs(sr.numpy() - z["sr"]).max() <= 1e-6
    assert abs(float(loss) - float(z["loss"])) <= 1e-7
    for k in z.files:
        if k.startswith("grad/"):
            g = grads[k[5:]].numpy()
            assert np.abs(g - z[k]).max() <= 1e-6 * max(1.0, np.abs(z[k]).max())


def test_state_dict_names_match_reference(golden_dir):
    z = _load(golden_dir, "gen_tiny_refinit.npz")
    ref_names = [k[3:] for k in z.files if k.startswith("sd/")]
    ours = synth.make_state_dict(2, 1, 64, 1, 16)
    assert list(ours.keys()) == ref_names
    for k in ref_names:
        assert tuple(ours[k].shape) == z["sd/" + k].shape


@pytest.mark.parametrize("name", ["gen_hydra_seeded.npz", "gen_default_seeded.npz"])
def test_seeded_configs_match_reference(golden_dir, name):
    z = _load(golden_dir, name)
    in_ch, nb, gc, n, h, w = (int(v) for v in z["meta"])
    x, elev, mask = synth.make_inputs(n, in_ch, h, w, seed=1)
    for tag, gain in zip(("sr", "sr_trained", "sr_stress"), (float(g) for g in z["gains"])):
        sd = synth.make_state_dict(in_ch, 1, 64, nb, gc, seed=0, gain=gain)
        with torch.no_grad():
            sr = og.generator_forward(sd, x, elev, mask)
        ref = z[tag]
        assert sr.shape == ref.shape
        assert np.abs(sr.numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("name", ["cfg2_default", "cfg2_trained", "cfg4_trained", "default64_trained"])
def test_fulldepth_configs_match_reference(golden_dir, name):
    """Full-depth outputs of the UNMODIFIED reference module at the BASELINE configs' own tile sizes (oracle/make_golden.py
    fulldepth()): two cfg2 tiles (64x64 LR, nb=11), the cfg4 Europe raster (113x113, in=3) and a class-default 64x64 tile."""
    from oracle import generator as og
    from oracle import synth
    z = np.load(os.path.join(golden_dir, "gen_fulldepth.npz"))
    in_ch, nb, gc, n, h, w, wseed, iseed = (int(v) for v in z[name + "_meta"])
    sd = synth.make_state_dict(in_ch, 1, 64, nb, gc, seed=wseed, gain=float(z[name + "_gain"]))
    x, elev, mask = synth.make_inputs(n, in_ch, h, w, seed=iseed, blocky_mask=name.startswith("cfg4"))
    with torch.no_grad():
        got = og.generator_forward(sd, x, elev, mask).numpy()
    assert got.shape == z[name].shape
    assert np.abs(got - z[name]).max() <= 2e-5 * max(1.0, np.abs(z[name]).max())


def test_numpy_restatement_agrees_with_torch_graph():
    sd = synth.make_state_dict(3, 1, 64, 1, 16, seed=3)
    x, elev, mask = synth.make_inputs(1, 3, 6, 5, seed=7)
    with torch.no_grad():
        a = og.generator_forward(sd, x, elev, mask).numpy()
    b = np_ops.generator_forward({k: v.numpy() for k, v in sd.items()}, x.numpy(), elev.numpy(), mask.numpy())
    assert a.shape == (1, 1, 24, 20)
    assert np.abs(a - b).max() <= 1e-5


def test_reference_shape_case():
    """tests/models/test_esrgan.py:7-22 of the reference (shape only), on a smaller batch for CPU time."""
    sd = synth.make_state_dict(2, 1, 64, 1, 16)
    x = torch.rand(2, 2, 32, 32)
    e = torch.rand(2, 1, 128, 128)
    m = torch.rand(2, 1, 128, 128)
    with torch.no_grad():
        assert og.generator_forward(sd, x, e, m).shape == (2, 1, 128, 128)


def test_flop_model_matches_survey():
    assert og.flops_per_hr_pixel(3, 64, 11, 16) == pytest.approx(721880.0)
    assert og.flops_per_hr_pixel(4, 64, 11, 16) == pytest.approx(721952.0)
    assert og.flops_per_hr_pixel(4, 64, 23, 32) == pytest.approx(2275424.0)


# ---- RegressionAccuracy known-answer cases: tests/metrics/test_regresion_accuracy.py:12-108 ----
SHAPE = (3, 128, 128)


def _kat_cases():
    for eps in (0.1, 1.0, 0.25):
        yield eps, torch.zeros(SHAPE), torch.ones(SHAPE) + (1 if eps == 1.0 else 0), 0.0
        yield eps, torch.ones(SHAPE), torch.ones(SHAPE), 1.0
        g = torch.Generator().manual_seed(int(eps * 100))
        yield eps, torch.ones(SHAPE) - torch.rand(SHAPE, generator=g) / 100, torch.ones(SHAPE), 1.0


@pytest.mark.parametrize("eps,preds,targets,expected", list(_kat_cases()))
def test_regression_accuracy_kats(eps, preds, targets, expected):
    assert float(om.regression_accuracy(preds, targets, eps)) == expected


def test_val_step_properties():
    g = torch.Generator().manual_seed(0)
    n, H, W = 2, 40, 36
    sr = torch.rand(n, 1, H, W, generator=g) * 2 - 1
    t = synth.make_targets(sr, seed=4)
    mask = (torch.rand(n, 1, H, W, generator=g) > 0.3).float()
    orig = om.denormalized_original(t["hr"], t["min"], t["max"])
    out = om.val_test_step(sr, t["hr"], orig, mask, t["min"], t["max"])
    assert set(om.ACC_KEYS) <= set(out)
    # identical inputs -> perfect scores
    same = om.val_test_step(t["hr"], t["hr"], orig, mask, t["min"], t["max"])
    assert float(same["mae"]) < 1e-5 and float(same["acc@0.1"]) == 1.0 and abs(float(same["ssim"]) - 1) < 1e-9
    # masked pixels do not contribute: perturb the ocean only
    sr2 = sr.clone()
    sr2[~mask.bool()] += 5.0
    out2 = om.val_test_step(sr2, t["hr"], orig, mask, t["min"], t["max"])
    for k in out:
        assert float(out[k]) == pytest.approx(float(out2[k]), rel=1e-12, abs=1e-12)
    # denominators are numel, not land count
    d = (om.minmax_denormalize(sr.double(), t["min"].double(), t["max"].double()) - orig.double()).abs() * mask
    assert float(out["mae"]) == pytest.approx(float(d.sum() / d.numel()), rel=1e-9)


def test_minmax_scaler_restatement_matches_reference_golden(golden_dir):
    """oracle/normalization.py against the reference's own MinMaxScaler (climsr/data/normalization.py:37-84, run unmodified
    by oracle/make_golden.py): normalize with NaN substitution, denormalize + NaN land mask (inference.py:73-76)."""
    from oracle import normalization as on
    g = _load(golden_dir, "normalization.npz")
    n = g["raw"].shape[0]
    norm = np.stack([on.normalize(g["raw"][i], g["mins"][i], g["maxes"][i]) for i in range(n)])
    assert np.array_equal(norm, g["norm"])
    assert not np.isnan(norm).any() and np.isnan(g["raw"]).any()
    assert np.array_equal(on.normalize(g["raw"][0], g["mins"][0], g["maxes"][0], (0.0, 1.0)), g["norm01"])
    post = on.postprocess(g["sr"], g["mask"], g["mins"], g["maxes"])
    assert np.array_equal(post, g["post"], equal_nan=True)
    assert np.array_equal(np.isnan(post[:, 0]), np.broadcast_to(g["mask"][0, 0] == 0, post[:, 0].shape))
    # round trip: denormalize(normalize(x)) == x on valid pixels (float32 rounding only)
    valid = ~np.isnan(g["raw"][1])
    back = on.denormalize(norm[1], g["mins"][1], g["maxes"][1])
    assert np.abs(back[valid] - g["raw"][1][valid]).max() <= 1e-4
    x = on.lr_input(g["raw"], g["mins"], g["maxes"], np.ones(g["raw"].shape[1:], np.float32), None)
    assert x.shape == (n, 2) + g["raw"].shape[1:] and np.array_equal(x[:, 0], norm)


def test_lr_input_restatement_matches_numpy_cv2_golden(golden_dir):
    """oracle/lr_input.py against numpy flips / rot90 + cv2.resize INTER_NEAREST (the arithmetic behind
    climate_dataset.py:152-172, run by oracle/make_golden.py): all 16 augmentation codes, exact."""
    from oracle import lr_input as ol
    g = _load(golden_dir, "lr_input.npz")
    x, hr, el, mk = ol.training_batch(g["hr"], g["elev"], g["mask"], g["codes"])
    assert np.array_equal(x, g["x"]) and np.array_equal(hr, g["hr_aug"]) and np.array_equal(el, g["elev_aug"]) and np.array_equal(mk, g["mask_aug"])
    assert np.array_equal(ol.resize_nearest(g["rect"]), g["rect_lr"])
    x0, hr0, _, _ = ol.training_batch(g["hr"], g["elev"], g["mask"], None)
    assert np.array_equal(hr0, g["hr"]) and np.array_equal(x0[:, 0], g["hr"][:, 0, ::4, ::4])
    # the four rotations compose to the identity; two flips cancel
    a = g["hr"][3, 0]
    assert np.array_equal(ol.augment(ol.augment(a, False, False, 1), False, False, 3), a)
    assert np.array_equal(ol.augment(ol.augment(a, True, True, 0), True, True, 0), a)


def test_discriminator_restatement_matches_reference_golden(golden_dir):
    """oracle/discriminator.py (groundwork for SURVEY 8f row 2) against the reference Discriminator run unmodified by
    oracle/make_golden.py on oracle.synth.make_discriminator_state_dict(0): train-mode scores (batch statistics), eval-mode
    scores, the relativistic losses of pl_gan.py:28-61 and two gradients of the discriminator loss."""
    from oracle import discriminator as od
    g = _load(golden_dir, "discriminator.npz")
    sd = synth.make_discriminator_state_dict(seed=0)
    assert list(sd.keys()) == [str(k) for k in g["names"]]
    for k in ("feature_extraction.1.weight", "classification.1.weight"):
        sd[k].requires_grad_(True)
    hr, sr = torch.from_numpy(g["hr"]), torch.from_numpy(g["sr"])
    s_real, s_fake = od.discriminator_forward(sd, hr), od.discriminator_forward(sd, sr)
    assert float((s_real.detach() - torch.from_numpy(g["s_real"])).abs().max()) <= 1e-5
    assert float((s_fake.detach() - torch.from_numpy(g["s_fake"])).abs().max()) <= 1e-5
    loss_g, loss_d = od.relativistic_losses(s_real, s_fake)
    assert abs(float(loss_g.detach()) - float(g["loss_g"])) <= 1e-6 and abs(float(loss_d.detach()) - float(g["loss_d"])) <= 1e-6
    g_first, g_last = torch.autograd.grad(loss_d, [sd["feature_extraction.1.weight"], sd["classification.1.weight"]])
    for got, want in ((g_first, g["g_first"]), (g_last, g["g_last"])):
        want = torch.from_numpy(want)
        assert float((got - want).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max()))
    with torch.no_grad():
        s_eval = od.discriminator_forward(sd, hr, training=False)
    assert float((s_eval - torch.from_numpy(g["s_eval"])).abs().max()) <= 1e-5


def _rcan_from_seed(cls, ng, nbk, seed):
    torch.manual_seed(seed)
    return cls(n_resgroups=ng, n_resblocks=nbk, n_feats=64, reduction=16, scaling_factor=4, in_channels=3, out_channels=1).eval()


def test_rcan_restatement_and_state_dict_contract_match_reference_golden(golden_dir):
    """oracle/rcan.py against the outputs of the UNMODIFIED reference RCAN (tests/golden/rcan.npz), with weights re-derived from
    the seed through climsr_b200.models.rcan.RCAN - which also pins that class's parameter names and creation order (= the
    torch.manual_seed initialisation contract) against the reference's state_dict."""
    from climsr_b200.models.rcan import RCAN
    from oracle import rcan as orc
    from oracle import synth
    z = np.load(os.path.join(golden_dir, "rcan.npz"))
    for tag in ("small", "hydra"):
        ng, nbk, n, h, w, seed = (int(v) for v in z[tag + "_meta"])
        net = _rcan_from_seed(RCAN, ng, nbk, seed)
        sd = net.state_dict()
        if tag == "small":
            assert list(sd.keys()) == [str(k) for k in z["names"]]
            assert np.allclose([float(v.double().sum()) for v in sd.values()], z["sd_sum"], rtol=0, atol=1e-9)
            assert np.allclose([float(v.double().abs().sum()) for v in sd.values()], z["sd_abs"], rtol=0, atol=1e-9)
        x, elev, mask = synth.make_inputs(n, 3, h, w, seed=50 + seed)
        with torch.no_grad():
            got = orc.rcan_forward(sd, x, elev, mask, ng, nbk).numpy()
        assert got.shape == z[tag].shape
        assert np.abs(got - z[tag]).max() <= 2e-5


def test_rcan_oracle_gradients_match_reference_autograd_golden(golden_dir):
    """Groundwork for RCAN training (the CUDA path is inference-only so far): autograd through oracle/rcan.py
    reproduces the gradients of the UNMODIFIED reference module (tests/golden/rcan_grad.npz, oracle/make_golden.py --rcan-grad-only)
    for every parameter (sum / abs-sum), four full tensors and the input."""
    from climsr_b200.models.rcan import RCAN
    from oracle import rcan as orc
    from oracle import synth
    z = np.load(os.path.join(golden_dir, "rcan_grad.npz"))
    ng, nbk, n, h, w, seed = (int(v) for v in z["meta"])
    net = _rcan_from_seed(RCAN, ng, nbk, seed)
    sd = {k: v.clone().requires_grad_(True) for k, v in net.state_dict().items()}
    x, elev, mask = synth.make_inputs(n, 3, h, w, seed=50 + seed)
    x.requires_grad_(True)
    wsum = torch.randn((n, 1, 4 * h, 4 * w), generator=torch.Generator().manual_seed(77))
    (orc.rcan_forward(sd, x, elev, mask, ng, nbk) * wsum).sum().backward()
    names = [str(k) for k in z["names"]]
    assert names == [k for k in sd.keys()]
    for i, k in enumerate(names):
        g = sd[k].grad
        assert g is not None, k
        assert abs(float(g.double().sum()) - float(z["g_sum"][i])) <= 1e-4 * max(1.0, float(z["g_abs"][i])), k
        assert abs(float(g.double().abs().sum()) - float(z["g_abs"][i])) <= 1e-4 * max(1.0, float(z["g_abs"][i])), k
    for key in z.files:
        if key.startswith("g:"):
            ref = z[key]
            assert np.abs(sd[key[2:]].grad.numpy() - ref).max() <= 1e-4 * max(1e-3, float(np.abs(ref).max())), key
    assert np.abs(x.grad.numpy() - z["dx"]).max() <= 1e-4 * max(1e-3, float(np.abs(z["dx"]).max()))
