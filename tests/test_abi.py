"""CPU: the C-ABI shared library builds, loads and exports every symbol include/hhfm_sm100.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from hhfm_b200 import build
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hhfm_sm100.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hhfm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_families():
    names = declared_symbols()
    for family in ("hhfm_pack_", "hhfm_fm_", "hhfm_pairrank_", "hhfm_scatter_add_rows", "hhfm_opt_", "hhfm_topn_",
                   "hhfm_metrics_walk", "hhfm_last_error"):
        assert any(n.startswith(family) for n in names), family


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, "declared in the header but not exported: %s" % missing


def test_library_exports_nothing_the_header_does_not_declare(lib_path):
    """Every `hhfm_*` symbol in the dynamic symbol table is part of the documented C ABI (no private entry points)."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = sorted({ln.split()[-1] for ln in out.splitlines() if re.search(r"\s[TW]\shhfm_[a-z0-9_]+$", ln)})
    assert exported, "nm found no hhfm_* exports"
    extra = sorted(set(exported) - set(declared_symbols()))
    assert not extra, "exported but not declared in include/hhfm_sm100.h: %s" % extra


def test_python_binding_covers_the_header(lib_path):
    from hhfm_b200 import _lib
    _lib.load()
    bound = set(_lib.exported_symbols())
    assert set(declared_symbols()) <= bound, sorted(set(declared_symbols()) - bound)


def test_bad_arguments_fail_loudly_without_a_gpu(lib_path):
    """Argument validation happens before any launch, so it is checkable on CPU."""
    from hhfm_b200 import _lib
    lib = _lib.load()
    assert lib.hhfm_abi_version() == 1
    assert lib.hhfm_partials_len() >= 1184
    with pytest.raises(_lib.HhfmError, match="K=6"):
        _lib.call("hhfm_fm_fwd", None, ctypes.c_void_p(16), None, 4, 3, ctypes.c_void_p(16), None, None, 10, 6, 0,
                  ctypes.c_void_p(16), None)
    with pytest.raises(_lib.HhfmError, match="n_neg"):
        _lib.call("hhfm_pairrank_fwd", ctypes.c_void_p(16), 4, 72, 2, 0, 65, 0, 0, 0, ctypes.c_void_p(16), 10, 8,
                  ctypes.c_void_p(16), None, None)
    p16 = ctypes.c_void_p(16)
    # evaluators / Wide&Deep / sparse-exchange helpers: shape checks run before any launch
    assert lib.hhfm_afm_topn_supported(10, 64, 64) == 1 and lib.hhfm_afm_topn_supported(10, 128, 128) == 0
    with pytest.raises(_lib.HhfmError, match="not covered"):
        _lib.call("hhfm_afm_topn_scores", p16, 12, 4, 10, 1, p16, None, None, p16, p16, p16, p16, 100, 128, 128, 10, 50, p16, p16, None)
    with pytest.raises(_lib.HhfmError, match="item_col"):
        _lib.call("hhfm_afm_topn_scores", p16, 12, 4, 10, 10, p16, None, None, p16, p16, p16, p16, 100, 64, 64, 10, 50, p16, p16, None)
    with pytest.raises(_lib.HhfmError, match="bad sizes"):
        _lib.call("hhfm_wd_wide_fwd", p16, 4, 33, p16, p16, p16, 100, 97, p16, None)
    with pytest.raises(_lib.HhfmError, match="bad argument"):
        _lib.call("hhfm_opt_ftrl_dense", p16, p16, p16, p16, 8, 0.0, 0.0, 0.0, 1, None)
    with pytest.raises(_lib.HhfmError, match="multiple of 4"):
        _lib.call("hhfm_gather_rows", p16, p16, 4, 6, 100, p16, 0, None)


def test_host_packer_narrows_and_validates(lib_path):
    import numpy as np
    from hhfm_b200 import _lib
    from hhfm_b200.engine import ptr
    src = np.arange(12, dtype=np.int64).reshape(4, 3)
    dst = np.full((4, 8), -7, dtype=np.int32)
    _lib.call("hhfm_pack_ids_i64", ptr(src), 4, 3, 3, ptr(dst), 8, 2, 12, 2)
    assert (dst[:, 2:5] == src).all() and (dst[:, :2] == -7).all() and (dst[:, 5:] == -7).all()
    _lib.call("hhfm_pack_fill_i32", ptr(dst), 4, 3, 8, 5, -1, 1)
    assert (dst[:, 5:] == -1).all()
    with pytest.raises(_lib.HhfmError, match="out of range"):
        _lib.call("hhfm_pack_ids_i64", ptr(src), 4, 3, 3, ptr(dst), 8, 2, 11, 2)
    rp = np.zeros(5, np.int32); col = np.zeros(12, np.int32)
    _lib.call("hhfm_pack_csr_i64", ptr(src), None, 4, 3, 3, ptr(rp), ptr(col), None, 12, 1)
    assert rp.tolist() == [0, 3, 6, 9, 12] and (col == np.arange(12)).all()


def test_missing_library_fails_loudly(monkeypatch):
    import pytest
    from hhfm_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libhhfm_sm100.so")
    with pytest.raises(_lib.HhfmError, match="no CPU fallback"):
        _lib.load()
