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


def assert_close(a, b, rtol=1e-5, what="", atol=0.0):
    """|a-b| <= rtol * max(|b|, rms(b)) element-wise: the 1e-5 relative fp32 tolerance of BASELINE.json with a
    floor at the RMS magnitude of the reference values (cancellation makes single elements arbitrarily small)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, "%s shape %s vs %s" % (what, a.shape, b.shape)
    if b.size == 0:
        return
    rms = float(np.sqrt(np.mean(b * b)))
    tol = rtol * np.maximum(np.abs(b), max(rms, 1e-30)) + atol
    err = np.abs(a - b)
    bad = err > tol
    assert not bad.any(), "%s: %d/%d elements off, worst err %.3e (tol %.3e) at %s" % (
        what, int(bad.sum()), b.size, float(err.max()), float(tol.reshape(-1)[np.argmax(err)]), np.unravel_index(np.argmax(err), b.shape))


def assert_update_close(w_got, w_ref, w_prev, g, acc_prev, lr, rtol=1e-5, what="", atol=0.0):
    """Post-step weights of one TF1-Adagrad step, compared as UPDATES (delta = w_new - w_prev) with a first-order
    propagation of the 1e-5 relative gradient tolerance through update(g) = lr*g/sqrt(acc+g^2):
        tol = rtol * ( max(|delta|, rms(delta)) + |d update/d g| * max(|g|, rms(g)) ),  d update/d g = lr*acc/(acc+g^2)^1.5.
    The second term only matters for the reference's acc0 = 1e-8 configs (BPR.py:93, MF.py:104), where the update
    of an element with |g| << 1e-4 is 100*g: a gradient that differs in the last fp32 bit moves the weight by more
    than 1e-5 of the typical update.  TF itself has the same conditioning."""
    w_got = np.asarray(w_got, np.float64); w_ref = np.asarray(w_ref, np.float64); w_prev = np.asarray(w_prev, np.float64)
    g = np.asarray(g, np.float64); acc_prev = np.asarray(acc_prev, np.float64)
    delta = w_ref - w_prev
    rms_d = float(np.sqrt(np.mean(delta * delta))); rms_g = float(np.sqrt(np.mean(g * g)))
    sens = lr * acc_prev / np.power(acc_prev + g * g, 1.5)
    tol = rtol * (np.maximum(np.abs(delta), rms_d) + sens * np.maximum(np.abs(g), rms_g)) + atol
    err = np.abs(w_got - w_ref)
    bad = err > tol
    assert not bad.any(), "%s: %d/%d elements off, worst err %.3e vs tol %.3e" % (
        what, int(bad.sum()), err.size, float(err.max()), float(tol.reshape(-1)[np.argmax(err)]))


@pytest.fixture(scope="session")
def cuda():
    import torch
    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    from hhfm_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")
