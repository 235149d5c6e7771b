import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly (not skip) when selected on a box without CUDA; they are deselected by -m 'not gpu'."""
    return


def assert_close(a, b, rtol=1e-5, what=""):
    """|a-b| <= rtol * max(|b|, rms(b)) element-wise: the 1e-5 relative fp32 tolerance of BASELINE.json with a
    floor at the RMS magnitude of the reference values (cancellation makes single elements arbitrarily small)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, "%s shape %s vs %s" % (what, a.shape, b.shape)
    if b.size == 0:
        return
    rms = float(np.sqrt(np.mean(b * b)))
    tol = rtol * np.maximum(np.abs(b), max(rms, 1e-30))
    err = np.abs(a - b)
    bad = err > tol
    assert not bad.any(), "%s: %d/%d elements off, worst err %.3e (tol %.3e) at %s" % (
        what, int(bad.sum()), b.size, float(err.max()), float(tol.reshape(-1)[np.argmax(err)]), np.unravel_index(np.argmax(err), b.shape))


@pytest.fixture(scope="session")
def cuda():
    import torch
    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    from hhfm_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")
