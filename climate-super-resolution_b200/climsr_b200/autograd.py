"""Training path of the generator (autograd through the CUDA forward).  Backward kernels: not built yet."""
from __future__ import annotations

from ._lib import CsrError


def generator_apply(module, x, elev, mask):
    raise CsrError("climsr_b200: generator backward (dgrad/wgrad kernels) is not implemented yet; "
                   "call the generator under torch.no_grad() / module.eval() with requires_grad_(False)")
