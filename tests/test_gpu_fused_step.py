"""GPU: the one-kernel tail of a training step (csrc/p2p.cu `dp_step_kernel`, n_ranks = 1 on the test box: hot-replica fold
+ optimizer + regulariser + loss in one cooperative launch) against (a) the three-kernel tail it replaces
(hhfm_hot_fold + hhfm_opt_*_dense_l2 + hhfm_loss_finalize, `HHFM_FUSED_STEP=0`), bit for bit, and (b) the oracle.
The cross-GPU phases of the same kernel are covered by tests/test_gpu_dist.py (>= 2 GPUs) and by the replica / NCCL
cross-check `bench.py` runs at N > 1."""
import numpy as np
import pytest
import torch

from conftest import assert_close, assert_update_close
from oracle import hhfm_oracle as O

pytestmark = pytest.mark.gpu


def _hhfm_batch(rng, n_user, n_item, M, fc, B):
    X = np.stack([rng.integers(0, n_user, B), n_user + rng.integers(0, n_item, B)], axis=1)
    F1 = rng.integers(n_user + n_item, M, (B, fc))
    Y = n_user + rng.integers(0, n_item, (B, 10))
    return {"X": X, "F1": F1, "Y": Y}


@pytest.mark.parametrize("opt", ["AdagradOptimizer", "AdamOptimizer", "GradientDescentOptimizer"])
@pytest.mark.parametrize("hot", [False, True])
def test_fused_tail_is_bit_identical_to_the_separate_kernels_hhfm(cuda, opt, hot):
    from hhfm_b200.models import OUR
    rng = np.random.default_rng(11)
    n_user, n_item, M, K, fc, B = 80, 150, 260, 64, 4, 3000
    lr = 0.1 if opt != "AdamOptimizer" else 0.01
    batches = [_hhfm_batch(rng, n_user, n_item, M, fc, B) for _ in range(4)]
    out = {}
    for fused in (True, False):
        m = OUR(fc, 0, M, n_user, n_item, K, lr, 0.01, opt, True, False)
        m._fused = fused
        m.hot_rows = list(range(n_user + n_item, M)) if hot else None       # the context rows: a few dozen hits each
        m.deterministic = not hot          # program-order scatter: the gradient itself is bit-reproducible (no replicas then)
        assert m._use_fused_tail() == fused
        losses = [m.partial_fit(b) for b in batches]
        out[fused] = (m.get_weights()["feature_embeddings"], losses, m._opt.state["feature_embeddings"][0].cpu().numpy()
                      if m._opt.kind != "sgd" else None)
    if not hot:
        assert np.array_equal(out[True][0], out[False][0]), "weights differ between the fused and the separate tail"
        if out[True][2] is not None:
            assert np.array_equal(out[True][2], out[False][2]), "optimizer state differs"
    else:
        # with replicas the scatter order is not reproducible between two runs: compare within the gradient tolerance
        # (four optimizer steps amplify last-bit differences of the gradient sums)
        assert_close(out[True][0], out[False][0], rtol=1e-3, what="weights (hot replicas)")
    assert_close(np.asarray(out[True][1]), np.asarray(out[False][1]), rtol=2e-6 if not hot else 1e-5, what="loss")


@pytest.mark.parametrize("K", [16, 64, 128])
def test_fused_tail_fm_segments_match_oracle(cuda, K):
    """FM: three segments (V with lamda, feature_bias, the scalar bias at an offset that is not a multiple of four)."""
    from hhfm_b200.models import FM
    rng = np.random.default_rng(3)
    n_user, n_item, M, F, B = 50, 77, 201, 6, 2048                     # M odd: the scalar bias gradient sits at n_v + 201
    X = np.concatenate([np.stack([rng.integers(0, n_user, B), n_user + rng.integers(0, n_item, B)], axis=1),
                        rng.integers(n_user + n_item, M, (B, F - 2))], axis=1)
    Y = rng.choice([1.0, 0.0], (B, 1)).astype(np.float32)
    res = {}
    for fused in (True, False):
        m = FM(F, M, n_user, n_item, K, 0.1, 0.1, 1, 'AdagradOptimizer', 0, 0)
        m._fused = fused
        m.deterministic = True             # program-order scatter: both tails see bit-identical gradients
        w0 = m.get_weights()
        loss = m.partial_fit({"X": X, "Y": Y})
        res[fused] = (m.get_weights(), loss)
        if fused:
            loss_ref, _, dV, db, db0, _ = O.fm_loss_grads(X, Y, w0["feature_embeddings"], w0["feature_bias"], w0["bias"], 0.1)
            V1, _ = O.adagrad_dense(w0["feature_embeddings"], np.full_like(w0["feature_embeddings"], 0.1), dV, 0.1)
            b1, _ = O.adagrad_dense(w0["feature_bias"], np.full_like(w0["feature_bias"], 0.1), np.asarray(db).reshape(-1, 1), 0.1)
            assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref)
            assert_update_close(res[True][0]["feature_embeddings"], V1, w0["feature_embeddings"], dV,
                                np.full_like(V1, 0.1), 0.1, what="V")
            assert_update_close(res[True][0]["feature_bias"], b1, w0["feature_bias"], np.asarray(db).reshape(-1, 1),
                                np.full_like(b1, 0.1), 0.1, what="feature_bias")
    for k in ("feature_embeddings", "feature_bias", "bias"):
        assert np.array_equal(res[True][0][k], res[False][0][k]), k
    assert abs(res[True][1] - res[False][1]) <= 2e-6 * abs(res[False][1])


def test_fused_tail_leaves_the_arena_clean_and_counts_steps(cuda):
    from hhfm_b200.models import BPR
    rng = np.random.default_rng(5)
    n_user, n_item, K, B = 64, 200, 128, 1024
    M = n_user + n_item
    m = BPR(M, n_user, n_item, K, 0.01, 0.1, 'AdagradOptimizer')
    assert m._use_fused_tail()
    for i in range(3):
        X = np.stack([rng.integers(0, n_user, B), n_user + rng.integers(0, n_item, B)], axis=1)
        m.partial_fit({"X": X, "Y": n_user + rng.integers(0, n_item, (B, 10))})
    assert float(m._arena[:M * K + 4].abs().max()) == 0.0
    assert m._dp_state.cpu().tolist()[:2] == [3, 0]


# ----------------------------------------------------------------------------------------------------
# dropout on the interaction layer (FM.py:114; MF.py:87 -- the reference's MF default is keep = 0.7)
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K,keep", [(64, 0.7), (128, 0.5), (100, 0.9)])
def test_mf_dropout_step_matches_oracle(cuda, K, keep):
    from hhfm_b200.models import MF
    rng = np.random.default_rng(21)
    n_user, n_item, B = 70, 130, 2500
    M = n_user + n_item
    m = MF(M, n_user, n_item, K, 0.01, 0.01, keep, 'AdagradOptimizer', 0, 0)
    V0 = m.get_weights()["feature_embeddings"].copy()
    X = np.stack([rng.integers(0, n_user, B), n_user + rng.integers(0, n_item, B)], axis=1)
    Y = rng.choice([1.0, -1.0], (B, 1)).astype(np.float32)
    loss = m.partial_fit({"X": X, "Y": Y})
    loss_ref, _, dV = O.mf_dropout_loss_grads(X, Y, V0, keep, m._last_drop_seed, 0.01)
    V1, _ = O.adagrad_dense(V0, np.full_like(V0, 1e-8), dV, 0.01)
    assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref), (loss, loss_ref)
    assert_update_close(m.get_weights()["feature_embeddings"], V1, V0, dV, np.full_like(V0, 1e-8), 0.01, what="MF dropout V")
    mask = O.dropout_mask_hashed(m._last_drop_seed, B, K, keep)
    assert abs(float(mask.mean()) - keep) < 0.01                                  # it is a Bernoulli(keep) mask
    # evaluation runs without dropout (MF.py:230: dropout_keep 1.0)
    out = m.predict(X[:64])
    want, _, _ = O.mf_forward(X[:64], m.get_weights()["feature_embeddings"])
    assert_close(out.reshape(-1), want, what="MF predict")


def test_fm_dropout_step_matches_oracle(cuda):
    from hhfm_b200.models import FM
    rng = np.random.default_rng(22)
    n_user, n_item, M, F, K, B, keep = 40, 60, 160, 5, 64, 2000, 0.8
    m = FM(F, M, n_user, n_item, K, 0.1, 0.1, keep, 'AdagradOptimizer', 0, 0)
    w0 = m.get_weights()
    X = np.concatenate([np.stack([rng.integers(0, n_user, B), n_user + rng.integers(0, n_item, B)], axis=1),
                        rng.integers(n_user + n_item, M, (B, F - 2))], axis=1)
    Y = rng.choice([1.0, 0.0], (B, 1)).astype(np.float32)
    loss = m.partial_fit({"X": X, "Y": Y})
    loss_ref, _, dV, db, db0 = O.fm_dropout_loss_grads(X, Y, w0["feature_embeddings"], w0["feature_bias"], w0["bias"], keep,
                                                       m._last_drop_seed, 0.1)
    V1, _ = O.adagrad_dense(w0["feature_embeddings"], np.full_like(w0["feature_embeddings"], 0.1), dV, 0.1)
    b1, _ = O.adagrad_dense(w0["feature_bias"], np.full_like(w0["feature_bias"], 0.1), np.asarray(db).reshape(-1, 1), 0.1)
    got = m.get_weights()
    assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref), (loss, loss_ref)
    assert_update_close(got["feature_embeddings"], V1, w0["feature_embeddings"], dV, np.full_like(V1, 0.1), 0.1, what="FM dropout V")
    assert_update_close(got["feature_bias"], b1, w0["feature_bias"], np.asarray(db).reshape(-1, 1), np.full_like(b1, 0.1), 0.1,
                        what="FM dropout bias")


def test_mf_dropin_main_runs_with_the_reference_defaults(cuda, tmp_path, monkeypatch):
    """Newcode/MF.py: keep = 0.7, lr = 0.01, acc0 = 1e-8, top-100 retrieval (MF.py:17-41,104,147)."""
    import os
    from test_gpu_train_dropin import _write_dataset
    from hhfm_b200.Newcode.MF import MF_main
    path = _write_dataset(str(tmp_path))
    monkeypatch.setenv("HHFM_RESULT_FILE", os.path.join(str(tmp_path), "result.txt"))
    np.random.seed(3)
    sess = MF_main("frappe", 32, argv=["--path", path, "--epoch", "6", "--verbose", "5", "--batch_size", "2048"])
    assert sess.model.keep == 0.7 and len(sess.loss_epoch) == 5 and all(np.isfinite(sess.loss_epoch))
    assert sess.loss_epoch[-1] < sess.loss_epoch[0]
    ids = sess.model.topk(np.asarray(sess.data.Test_data.values[:20, 1:], dtype=np.int64))
    assert ids.shape == (20, 100) and ids.min() >= 0 and ids.max() < sess.n_item


def test_fused_tail_folds_bias_replicas(cuda):
    """FM with hot-row replicas (embedding and feature_bias replicas folded inside the one-kernel tail) against the oracle."""
    from hhfm_b200.models import FM
    rng = np.random.default_rng(9)
    n_user, n_item, M, F, K, B = 30, 50, 120, 5, 64, 4000
    X = np.concatenate([np.stack([rng.integers(0, n_user, B), n_user + rng.integers(0, n_item, B)], axis=1),
                        rng.integers(n_user + n_item, M, (B, F - 2))], axis=1)
    Y = rng.choice([1.0, 0.0], (B, 1)).astype(np.float32)
    m = FM(F, M, n_user, n_item, K, 0.1, 0.1, 1, 'AdagradOptimizer', 0, 0)
    m.hot_rows = list(range(0, 10)) + list(range(n_user + n_item, M))
    assert m._use_fused_tail()
    w0 = m.get_weights()
    loss = m.partial_fit({"X": X, "Y": Y})
    assert m._hot is not None and m._hot.ghot_bias is not None
    loss_ref, _, dV, db, db0, _ = O.fm_loss_grads(X, Y, w0["feature_embeddings"], w0["feature_bias"], w0["bias"], 0.1)
    V1, _ = O.adagrad_dense(w0["feature_embeddings"], np.full_like(w0["feature_embeddings"], 0.1), dV, 0.1)
    b1, _ = O.adagrad_dense(w0["feature_bias"], np.full_like(w0["feature_bias"], 0.1), np.asarray(db).reshape(-1, 1), 0.1)
    got = m.get_weights()
    assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref)
    assert_update_close(got["feature_embeddings"], V1, w0["feature_embeddings"], dV, np.full_like(V1, 0.1), 0.1, what="V")
    assert_update_close(got["feature_bias"], b1, w0["feature_bias"], np.asarray(db).reshape(-1, 1), np.full_like(b1, 0.1), 0.1, what="bias")
    assert float(m._hot.ghot.abs().max()) == 0.0 and float(m._hot.ghot_bias.abs().max()) == 0.0


# ----------------------------------------------------------------------------------------------------
# lazy-exact dense L2 (hard part 2-ii): untouched rows replay their `g = lamda * w` steps when they are next
# gathered / at flush -- bit-identical to the dense update of every row every step
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("which", ["OUR", "FM", "BPR"])
def test_lazy_l2_replay_is_bit_identical_to_the_dense_update(cuda, which):
    from hhfm_b200.models import BPR, FM, OUR
    rng = np.random.default_rng(17)
    n_user, n_item, M, K, fc, B, steps = 300, 900, 1400, 64, 4, 96, 50
    res = {}
    for lazy in (True, False):
        if which == "OUR":
            m = OUR(fc, 0, M, n_user, n_item, K, 0.1, 0.01, 'AdagradOptimizer', True, False)
        elif which == "BPR":
            m = BPR(M, n_user, n_item, K, 0.05, 0.1, 'AdagradOptimizer')
        else:
            m = FM(2 + fc, M, n_user, n_item, K, 0.1, 0.1, 1, 'AdagradOptimizer', 0, 0)
        m.lazy_l2 = lazy
        m.deterministic = True             # program-order scatter: the data gradient is bit-reproducible
        assert m._lazy() == lazy
        r2 = np.random.default_rng(5)
        mid = None
        for step in range(steps):
            X = np.stack([r2.integers(0, n_user, B), n_user + r2.integers(0, n_item, B)], axis=1)
            F1 = r2.integers(n_user + n_item, M - 50, (B, fc))          # the last 50 rows are never gathered
            if which == "FM":
                m.partial_fit({"X": np.concatenate([X, F1], axis=1), "Y": r2.choice([1.0, 0.0], (B, 1)).astype(np.float32)})
            elif which == "BPR":
                m.partial_fit({"X": X, "Y": n_user + r2.integers(0, n_item, (B, 10))})
            else:
                m.partial_fit({"X": X, "F1": F1, "Y": n_user + r2.integers(0, n_item, (B, 10))})
            if step == 20:                  # a read in the middle of training flushes: the lists must agree as well
                A = np.concatenate([X[:32], F1[:32]], axis=1) if which != "BPR" else X[:32]
                mid = m.topk(A, 20)
        w = m.get_weights()["feature_embeddings"]                        # flushes
        acc = m._opt.state["feature_embeddings"][0].cpu().numpy()
        res[lazy] = (w, acc, mid)
    assert np.array_equal(res[True][0], res[False][0]), "weights differ between the lazy replay and the dense update"
    assert np.array_equal(res[True][1], res[False][1]), "Adagrad accumulators differ"
    assert np.array_equal(res[True][2], res[False][2]), "top-N lists after the mid-training flush differ"
    w0 = np.asarray(res[True][0])
    assert not np.array_equal(w0[-50:], np.zeros_like(w0[-50:]))


def test_partial_fit_async_returns_the_same_losses_with_one_step_in_flight(cuda):
    """`partial_fit_async` + `PendingLoss.result()` (the trainers' pipelined epoch loop) against the blocking call: the same
    floats in the same order, also when the four-slot result ring wraps."""
    from hhfm_b200.models import OUR
    from hhfm_b200.trainer import _PipelinedFit
    rng = np.random.default_rng(21)
    n_user, n_item, M, K, fc, B = 80, 150, 260, 64, 4, 1500
    batches = [_hhfm_batch(rng, n_user, n_item, M, fc, B) for _ in range(7)]
    out = {}
    for mode in ("blocking", "async"):
        m = OUR(fc, 0, M, n_user, n_item, K, 0.1, 0.01, "AdagradOptimizer", True, False)
        m.deterministic = True
        if mode == "blocking":
            losses = [m.partial_fit(b) for b in batches]
        else:
            losses, prev = [], None
            for b in batches:
                h = m.partial_fit_async(b)
                if prev is not None:
                    losses.append(prev.result())
                prev = h
            losses.append(prev.result())
        out[mode] = (losses, m.get_weights()["feature_embeddings"])
    assert out["blocking"][0] == out["async"][0]
    assert np.array_equal(out["blocking"][1], out["async"][1])
    m = OUR(fc, 0, M, n_user, n_item, K, 0.1, 0.01, "AdagradOptimizer", True, False)
    m.deterministic = True
    fit = _PipelinedFit(m)
    for b in batches:
        fit(b)
    assert abs(fit.total() - sum(out["blocking"][0])) <= 1e-6 * abs(sum(out["blocking"][0]))
