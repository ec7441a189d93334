"""climsr_b200 - B200-native (sm_100a) generator hot path of xultaeculcis/climate-super-resolution.

Host-side mirror of the reference interface for that path:
  climsr_b200.models.esrgan.ESRGANGenerator   <- climsr/models/esrgan.py:57-102 (drop-in nn.Module, same state_dict)
  climsr_b200.metrics.masked_val_metrics      <- climsr/core/task.py:262-300,342-380
  climsr_b200.tiling                          <- halo-tiled multi-GPU raster inference
All arithmetic runs in hand-written CUDA behind the C-ABI of include/climsr_b200.h (libclimsr_b200.so, built
in-tree by build.py).  There is no CPU / eager fallback: importing works anywhere, computing needs a B200.
"""
from ._lib import lib, CsrError, NetDesc, device_check, kernel_launch_count  # noqa: F401

__all__ = ["lib", "CsrError", "NetDesc", "device_check", "kernel_launch_count"]
