"""CPU-only host-logic tests of the drop-in module (no kernels run)."""
import os

import numpy as np
import pytest
import torch

from climsr_b200.models import ESRGANGenerator


def test_state_dict_matches_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "gen_tiny_refinit.npz"))
    torch.manual_seed(0)
    net = ESRGANGenerator(in_channels=2, out_channels=1, nf=64, nb=1, gc=16)
    sd = net.state_dict()
    ref_names = [k[3:] for k in z.files if k.startswith("sd/")]
    assert list(sd.keys()) == ref_names
    # same creation order + same seed -> bit-identical default init as the reference ctor (esrgan.py:72-87)
    for k in ref_names:
        assert np.array_equal(sd[k].numpy(), z["sd/" + k]), k


def test_ctor_signature_and_kwargs():
    net = ESRGANGenerator(in_channels=3, out_channels=1, nf=64, nb=2, gc=16, scaling_factor=4, scale_factor=4, foo="bar")
    assert net.scale_factor == 4
    assert sum(p.numel() for p in net.parameters()) > 0
    with pytest.raises(ValueError):
        ESRGANGenerator(in_channels=3, out_channels=3)      # SRCNN tail only ever worked with 1 (esrgan.py:87,100)
    with pytest.raises(ValueError):
        ESRGANGenerator(in_channels=3, out_channels=1, scaling_factor=2)


def test_param_counts_match_survey():
    count = lambda m: sum(p.numel() for p in m.parameters())  # noqa: E731
    assert count(ESRGANGenerator(3, 1, 64, 11, 16)) == 4_278_530
    assert count(ESRGANGenerator(4, 1, 64, 11, 16)) == 4_279_106
    assert count(ESRGANGenerator(4, 1, 64, 23, 32)) == 16_715_906


def test_forward_validates_shapes():
    net = ESRGANGenerator(2, 1, 64, 1, 16).eval()
    with torch.no_grad():
        with pytest.raises(ValueError):
            net(torch.rand(1, 3, 8, 8), torch.rand(1, 1, 32, 32), torch.rand(1, 1, 32, 32))
        with pytest.raises(ValueError):
            net(torch.rand(1, 2, 8, 8), torch.rand(1, 1, 16, 16), torch.rand(1, 1, 32, 32))


def test_load_state_dict_from_oracle_names():
    from oracle import synth
    sd = synth.make_state_dict(4, 1, 64, 2, 16, seed=3)
    net = ESRGANGenerator(4, 1, 64, 2, 16)
    missing, unexpected = net.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    # Lightning checkpoints prefix the generator with "generator." and are loaded with strict=False (inference.py:125)
    pref = {"generator." + k: v for k, v in sd.items()}
    holder = torch.nn.Module()
    holder.generator = ESRGANGenerator(4, 1, 64, 2, 16)
    r = holder.load_state_dict(pref, strict=False)
    assert not r.missing_keys and not r.unexpected_keys
