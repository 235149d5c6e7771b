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
        assert m._use_fused_tail() == fused
        losses = [m.partial_fit(b) for b in batches]
        out[fused] = (m.get_weights()["feature_embeddings"], losses, m._opt.state["feature_embeddings"][0].cpu().numpy()
                      if m._opt.kind != "sgd" else None)
    assert np.array_equal(out[True][0], out[False][0]), "weights differ between the fused and the separate tail"
    if out[True][2] is not None:
        assert np.array_equal(out[True][2], out[False][2]), "optimizer state differs"
    assert_close(np.asarray(out[True][1]), np.asarray(out[False][1]), rtol=2e-6, what="loss")


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
        m.hot_rows = [0, 1, n_user, M - 1]
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
    assert float(m._arena[:M * K + M + 4].abs().max()) == 0.0
    assert m._dp_state.cpu().tolist()[:2] == [3, 0]
