import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu() -> bool:
    try:
        from frave_b200 import capi
        return capi.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# ---- synthetic inputs of SURVEY.md §8(d) -------------------------------------------------
def uniform_image(h, w, c, seed, dtype=np.uint8):
    """U: i.i.d. uniform over the full sample range."""
    rng = np.random.Generator(np.random.PCG64(seed))
    hi = np.iinfo(dtype).max
    return rng.integers(0, hi + 1, size=(h, w, c), dtype=dtype)


def smooth_image(h, w, c, seed, dtype=np.uint8):
    """S: smooth sinusoid + small noise (keeps residuals small)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    hi = np.iinfo(dtype).max
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    out = np.empty((h, w, c), dtype)
    for ch in range(c):
        v = 0.5 + 0.375 * np.sin(2 * np.pi * x / w * 3 + ch) * np.cos(2 * np.pi * y / h * 2 + 0.5 * ch)
        v = v * hi + rng.normal(0, 2, size=(h, w))
        out[:, :, ch] = np.clip(np.rint(v), 0, hi).astype(dtype)
    return out


def smallest_layer_q(d, both=True):
    """q[8] = q[9] = d (the 'smallest layer' of README.md:12), everything else 1."""
    q = np.ones(32, np.int32)
    q[8] = d
    if both:
        q[9] = d
    return q
