import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    """The C-ABI shared object (built in-tree if nvcc is around and the .so is stale)."""
    from volsurfs_b200 import _lib, build

    if not build.is_fresh():
        try:
            build.build()
        except Exception as exc:  # pragma: no cover
            if not build.LIB_PATH.exists():
                pytest.fail(f"cannot build libvolsurfs_b200.so: {exc}")
    return _lib.lib()


def rel_err(a, b, floor=1e-6):
    """max |a-b| / max(|b|, floor) — the metric of SURVEY.md section 8d (C1)."""
    import numpy as np

    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def grad_err(a, b):
    """Error metric for gradients (sums of mixed-sign terms, so individual entries can cancel to ~0):
    max |a-b| / max(|b|, rms(b)) — relative to the entry itself unless it is smaller than the tensor's typical
    magnitude, in which case relative to that magnitude.  The fp32 torch-autograd reference itself differs from the
    fp64 truth by ~1e-6 under this metric (tests/test_gpu_compositing.py::test_reference_fp32_error_scale)."""
    import numpy as np

    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    rms = float(np.sqrt(np.mean(b * b))) or 1.0
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), rms)))
