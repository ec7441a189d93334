"""CPU fp32 restatement of the reference discriminator and of the relativistic GAN losses.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Groundwork for SURVEY.md section 8f row 2 (the discriminator is NOT on the
CUDA path yet; tools/bench_gan_step.py runs it as stock PyTorch).  Pinned by tests/golden/discriminator.npz, generated from
the reference module itself by oracle/make_golden.py.

Follows climsr/models/discriminator.py:5-46: four stages of [ReflectionPad2d(1), Conv2d(cin, c, 3), LeakyReLU(0.01),
BatchNorm2d(c), ReflectionPad2d(1), Conv2d(c, c, 3, stride 2), LeakyReLU(0.01)] with c = 64, 128, 256, 512, then
Conv2d(512, 512, 3) (no padding), LeakyReLU(0.2), Conv2d(512, 512, 3), flatten, Linear(8192, 100), Linear(100, 1) - a
128 x 128 input gives 8 x 8 -> 6 x 6 -> 4 x 4 x 512 = 8192 features, the only size the module accepts (:40).  The
``avgpool`` member (:38) is never called.  State-dict names are those of the reference's nn.Sequential indices.
Losses: climsr/task/pl_gan.py:28-61.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def discriminator_forward(sd: Dict[str, Tensor], x: Tensor, training: bool = True, eps: float = 1e-5) -> Tensor:
    """``feature_extraction`` + ``classification`` (discriminator.py:42-46).  BatchNorm in training mode uses the batch
    statistics (biased variance), as ``nn.BatchNorm2d`` does under ``.train()``; running statistics are not updated here."""
    idx = 0
    for _ in range(4):
        x = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), sd[f"feature_extraction.{idx + 1}.weight"], sd[f"feature_extraction.{idx + 1}.bias"])
        x = F.leaky_relu(x, 0.01)
        bn = f"feature_extraction.{idx + 3}"
        if training:
            x = F.batch_norm(x, None, None, sd[bn + ".weight"], sd[bn + ".bias"], True, 0.0, eps)
        else:
            x = F.batch_norm(x, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"], False, 0.0, eps)
        x = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), sd[f"feature_extraction.{idx + 5}.weight"], sd[f"feature_extraction.{idx + 5}.bias"],
                     stride=2)
        x = F.leaky_relu(x, 0.01)
        idx += 7
    x = F.leaky_relu(F.conv2d(x, sd[f"feature_extraction.{idx}.weight"], sd[f"feature_extraction.{idx}.bias"]), 0.2)
    x = F.conv2d(x, sd[f"feature_extraction.{idx + 2}.weight"], sd[f"feature_extraction.{idx + 2}.bias"])
    x = x.reshape(x.shape[0], -1)
    x = F.linear(x, sd["classification.0.weight"], sd["classification.0.bias"])
    return F.linear(x, sd["classification.1.weight"], sd["classification.1.bias"])


def relativistic_losses(score_real: Tensor, score_fake: Tensor) -> Tuple[Tensor, Tensor]:
    """(adversarial loss of the generator, loss of the discriminator) from D(hr), D(sr): pl_gan.py:31-39 and :52-59."""
    ones, zeros = torch.ones_like(score_real), torch.zeros_like(score_real)
    rf = score_real - score_fake.mean()
    fr = score_fake - score_real.mean()
    bce = F.binary_cross_entropy_with_logits
    loss_g = (bce(fr, ones) + bce(rf, zeros)) / 2
    loss_d = (bce(fr, zeros) + bce(rf, ones)) / 2
    return loss_g, loss_d
