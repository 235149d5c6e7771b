"""GPU: single-touch rows (include/hhfm_sm100.h K14).  The staged scatter kernels apply the optimizer step of a row that
exactly one sample of the step references on the spot instead of sending its gradient through the arena and the rows
optimizer.  Both ways run the same element update, so weights, accumulators and loss must agree BIT FOR BIT -- on batches
in which no row takes more than two contributions (a + b is commutative; three or more atomically added terms are not
reproducible between two runs of the plain path either)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ids_at_most_twice(rng, n_rows, n_cols, M):
    """[n_rows, n_cols] ids in [0, M): half of the slots hold ids used once, the rest ids used twice."""
    T = n_rows * n_cols
    n_once, n_twice = T // 2, (T - T // 2) // 2
    assert n_once + n_twice <= M and n_once + 2 * n_twice == T
    perm = rng.permutation(M)
    ids = np.concatenate([perm[:n_once], np.repeat(perm[n_once:n_once + n_twice], 2)])
    rng.shuffle(ids)
    return ids.reshape(n_rows, n_cols).astype(np.int64)


def test_count_refs_matches_bincount(cuda):
    from hhfm_b200 import _lib
    from hhfm_b200.engine import cur_stream, ptr
    rng = np.random.default_rng(0)
    M, B, stride, cols = 1000, 777, 12, 10
    ids = rng.integers(0, M, (B, stride)).astype(np.int32)
    ids[rng.random((B, stride)) < 0.05] = -1                                   # padding
    t = torch.from_numpy(ids).to(cuda)
    cnt = torch.full((M,), 7, dtype=torch.int32, device=cuda)                 # must be cleared by the call
    _lib.call("hhfm_count_refs", ptr(t), B, stride, cols, M, ptr(cnt), cur_stream())
    sel = ids[:, :cols]
    want = np.bincount(sel[sel >= 0].ravel(), minlength=M)
    assert np.array_equal(cnt.cpu().numpy(), want)


@pytest.mark.parametrize("opt", ["AdagradOptimizer", "GradientDescentOptimizer"])
def test_fm_single_touch_rows_are_bit_identical_to_the_arena_path(cuda, opt, monkeypatch):
    from hhfm_b200.models import FM
    monkeypatch.setenv("HHFM_FM_STAGED", "1")            # the large-table kernel on a small table
    rng = np.random.default_rng(5)
    M, K, F, B = 40000, 64, 10, 2048
    batches = []
    for _ in range(3):
        X = torch.from_numpy(_ids_at_most_twice(rng, B, F, M).astype(np.int32)).to(cuda)
        y = torch.from_numpy(rng.choice([1.0, 0.0], B).astype(np.float32)).to(cuda)
        batches.append((X, y))
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("HHFM_SINGLE_TOUCH", mode)
        m = FM(F, M, 100, 100, K, 0.05 if opt == "AdagradOptimizer" else 0.002, 0.0, 1, opt, 0, 0)
        m.hot_rows = None
        snaps, touched = [], []
        for X, y in batches:
            m.fit_device(X, y)
            touched.append(int(m._touch.count.item()))
            w = m.get_weights()
            st = m._opt.state
            snaps.append((w["feature_embeddings"].copy(), w["feature_bias"].copy(), m._read_loss(),
                          st["feature_embeddings"][0].cpu().numpy().copy() if opt == "AdagradOptimizer" else None,
                          st["feature_bias"][0].cpu().numpy().copy() if opt == "AdagradOptimizer" else None))
        res[mode] = (snaps, touched)
    on, off = res["1"], res["0"]
    # step 1, from identical weights: the same bits
    a, b = on[0][0], off[0][0]
    assert np.array_equal(a[0], b[0]), "embedding table differs after one step"
    assert np.array_equal(a[1], b[1]), "feature_bias differs after one step"
    assert a[2] == b[2], "loss differs"
    if a[3] is not None:
        assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4]), "Adagrad accumulators differ after one step"
    # later steps: the scalar bias gradient is a sum of float atomics over the CTAs (order not reproducible between two runs of
    # EITHER path), so from step 2 on two runs agree to rounding, not to the bit
    a, b = on[0][-1], off[0][-1]
    np.testing.assert_allclose(a[0], b[0], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(a[1], b[1], rtol=1e-4, atol=1e-7)
    if a[3] is not None:
        np.testing.assert_allclose(a[3], b[3], rtol=1e-5)
    # the plan really took rows: the touched-row list of the plain path holds every distinct id of the batch, the other one
    # misses the rows that were updated in place (at most three per sample)
    for t_on, t_off in zip(on[1], off[1]):
        assert t_off == B * F // 2 + B * F // 4
        assert t_off - 3 * B <= t_on < t_off - B, (t_on, t_off)


@pytest.mark.parametrize("opt", ["AdagradOptimizer", "GradientDescentOptimizer"])
def test_hhfm_single_touch_rows_are_bit_identical_to_the_arena_path(cuda, opt, monkeypatch):
    from hhfm_b200.models import OUR
    monkeypatch.setenv("HHFM_PR_STAGED", "1")
    rng = np.random.default_rng(6)
    M, K, fc, NG, B = 60000, 64, 8, 10, 2048
    W = 2 + fc + NG
    recs = [torch.from_numpy(_ids_at_most_twice(rng, B, W, M).astype(np.int32)).to(cuda) for _ in range(3)]
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("HHFM_SINGLE_TOUCH", mode)
        m = OUR(fc, 0, M, 1000, 1000, K, 0.05, 0.0, opt, True, False)
        m.hot_rows = None
        losses, touched = [], []
        for r in recs:
            m.fit_device(r, fc, 0, NG)
            losses.append(m._read_loss())
            touched.append(int(m._touch.count.item()))
        res[mode] = (m.get_weights()["feature_embeddings"], losses, touched,
                     m._opt.state["feature_embeddings"][0].cpu().numpy() if opt == "AdagradOptimizer" else None)
    on, off = res["1"], res["0"]
    assert np.array_equal(on[0], off[0]), "embedding table differs"
    assert on[1] == off[1], "loss differs"
    if on[3] is not None:
        assert np.array_equal(on[3], off[3]), "Adagrad accumulator differs"
    for t_on, t_off in zip(on[2], off[2]):
        assert t_on < t_off - B // 2, (t_on, t_off)          # user / item+ / context rows referenced once went in place


def test_single_touch_plan_is_rejected_for_optimizers_that_move_untouched_rows(cuda):
    import ctypes as C
    from hhfm_b200 import _lib
    from hhfm_b200.engine import NO_HOT, SingleTouchPlan, cur_stream, ptr
    M, K, B = 100, 64, 8
    V = torch.zeros(M, K, device=cuda); g = torch.zeros(M, K, device=cuda); acc = torch.ones(M, K, device=cuda)
    idx = torch.zeros(B, 12, dtype=torch.int32, device=cuda)
    cnt = torch.zeros(M, dtype=torch.int32, device=cuda)
    lp = torch.zeros(_lib.partials_len(), device=cuda)
    stamp = torch.zeros(M, dtype=torch.int32, device=cuda); rows = torch.zeros(M, dtype=torch.int32, device=cuda)
    n = torch.zeros(1, dtype=torch.int32, device=cuda)
    plan = SingleTouchPlan()
    plan.ref_count = ptr(cnt); plan.V = ptr(V); plan.acc = ptr(acc); plan.lr = 0.1; plan.opt_kind = 1       # Adam
    with pytest.raises(_lib.HhfmError):
        _lib.call("hhfm_pairrank_fwd_bwd_st", ptr(idx), B, 12, 0, 0, 10, 0, 0, 0, ptr(V), M, K, None, None, ptr(g), ptr(lp),
                  ptr(stamp), 1, ptr(rows), ptr(n), *NO_HOT, 0, C.addressof(plan), cur_stream())
