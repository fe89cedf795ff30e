import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    # a fresh checkout has no built library (it is git-ignored): build it once, like the driver's build() step
    try:
        import __graft_entry__ as entry
        if not os.path.exists(entry.LIB) and os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):
            entry.build()
    except Exception as e:      # the tests that need the library will say so
        print(f"conftest: could not build the native library: {e}")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, f"{name}.npz"), allow_pickle=False))


def golden_data(fix):
    """{image index: dict(data=, variance=)} from an LM fixture."""
    out, i = {}, 0
    while f"data{i}" in fix:
        out[i] = {"data": fix[f"data{i}"], "variance": fix[f"var{i}"]}
        i += 1
    return out


def rel_err(a, b):
    """max |a-b| / max |b|  (image-scale relative error)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
