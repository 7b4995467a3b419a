"""volsurfs_b200 — B200-native (sm_100a) implementation of the per-ray rendering hot path of
autonomousvision/volsurfs: K-layer shell intersection -> RaySamplesPacked packing -> appearance MLP ->
alpha compositing forward/backward.  CUDA kernels behind a C ABI (``include/volsurfs_b200.h``), PyTorch for
device memory, streams and ``torch.distributed`` only.

Importing this package does not need a GPU; calling an operator does.  There is no CPU fallback.
"""
from __future__ import annotations

import sys

__version__ = "0.1.0"


def install_as_volsurfs() -> None:
    """Register :mod:`volsurfs_b200.volsurfs` as the top-level module ``volsurfs`` so the reference's
    ``from volsurfs import VolumeRendering, RaySamplesPacked`` (volume_rendering_funcs.py:5) resolves to
    the B200 kernels."""
    from . import volsurfs as _shim

    sys.modules["volsurfs"] = _shim
