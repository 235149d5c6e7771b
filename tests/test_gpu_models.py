"""GPU parity tests at the reference-facing API: model classes (`partial_fit`, `topk`, `sess.run`) against the
oracle's restatement of the TF graphs + TF1 optimizers, with injected weights (the reference is unseeded)."""
import os

import numpy as np
import pytest

from conftest import assert_close
from oracle import hhfm_oracle as O

pytestmark = pytest.mark.gpu


def frappe_like(rng, B, n_user=957, n_item=4082, ctx=(7, 2, 3, 2, 9, 80, 233, 7)):
    """Id rows [B, 2+len(ctx)] in the reference's column-major id layout (users, items, then context columns)."""
    cols = [rng.integers(0, n_user, B), n_user + rng.integers(0, n_item, B)]
    base = n_user + n_item
    for c in ctx:
        cols.append(base + rng.integers(0, c, B))
        base += c
    return np.stack(cols, axis=1), base


def test_fm_n_steps_match_oracle(cuda):
    from hhfm_b200.models import FM
    rng = np.random.default_rng(0)
    X0, M = frappe_like(rng, 5000)
    K, lam, lr = 64, 0.1, 0.1
    model = FM(10, M, 957, 4082, K, lr, lam, 1, 'AdagradOptimizer', 0, 0)
    w = model.get_weights()
    V = w["feature_embeddings"].copy(); b = rng.normal(0, 0.01, (M, 1)).astype(np.float32); b0 = np.float32(0.0)
    model.load_weights({"feature_bias": b})
    accV = np.full_like(V, 0.1); accb = np.full(M, 0.1, np.float32); accb0 = np.float32(0.1)
    for step in range(10):
        X, _ = frappe_like(rng, 5000 if step < 9 else 1234)
        Y = rng.choice([1.0, 0.0], (len(X), 1)).astype(np.float32)
        loss_ref, _, dV, db, db0, _ = O.fm_loss_grads(X, Y, V, b, b0, lam)
        V, accV = O.adagrad_dense(V, accV, dV, lr)
        bb, accb = O.adagrad_dense(b.reshape(-1), accb, db, lr); b = bb.reshape(-1, 1)
        b0, accb0 = O.adagrad_dense(b0, accb0, db0, lr)
        loss = model.partial_fit({"X": X, "Y": Y})
        assert_close(loss, loss_ref, what="loss step %d" % step)
    got = model.get_weights()
    assert_close(got["feature_embeddings"], V, what="V after 10 steps")
    assert_close(got["feature_bias"], b, what="bias after 10 steps")
    assert_close(got["bias"], b0, what="b0 after 10 steps")
    # forward-only entry used by evaluate_AUC: sess.run(model.out, feed_dict)
    out = model.sess.run(model.out, feed_dict={model.train_features: X0[:600], model.train_labels: [[1]] * 600,
                                               model.dropout_keep: 1.0, model.train_phase: False})
    assert out.shape == (600, 1)
    assert_close(out[:, 0], O.fm_forward(X0[:600], V, b, b0)[0], what="predict")
    # top-N through the API, bit-exact lists
    A = X0[:300]
    ids = model.topk(A, 20)
    Vd, bd = got["feature_embeddings"], got["feature_bias"]
    assert ids.dtype == np.int32 and ids.shape == (300, 20)
    assert (ids == O.topk_lowest_index(O.fm_topk_scores(A, Vd, bd, 957, 4082), 20)).all()


@pytest.mark.parametrize("opt", ["AdamOptimizer", "MomentumOptimizer", "GradientDescentOptimizer", "AdagradOptimizer"])
def test_fm_sparse_mode_optimizers(cuda, opt):
    """lamda = 0: IndexedSlices semantics (only touched rows move; TF1 sparse Adam moves every row)."""
    from hhfm_b200.models import FM
    rng = np.random.default_rng(1)
    M, K, lr = 300, 32, 0.01
    model = FM(6, M, 50, 100, K, lr, 0.0, 1, opt, 0, 0)
    V = model.get_weights()["feature_embeddings"].copy(); b = np.zeros(M, np.float32); b0 = np.float32(0)
    V_init = V.copy()
    s = {k: [np.zeros_like(V), np.zeros(M, np.float32), np.float32(0)] for k in ("m", "v")}
    if opt == "AdagradOptimizer":
        s["m"] = [np.full_like(V, 0.1), np.full(M, 0.1, np.float32), np.float32(0.1)]
    for t in range(1, 6):
        X = rng.integers(0, 150, (400, 6)); Y = rng.choice([1.0, -1.0], (400, 1)).astype(np.float32)
        loss_ref, _, dV, db, db0, rows = O.fm_loss_grads(X, Y, V, b.reshape(-1, 1), b0, 0.0)
        if opt == "AdamOptimizer":
            V, s["m"][0], s["v"][0] = O.adam_dense(V, s["m"][0], s["v"][0], dV, lr, t)
            b, s["m"][1], s["v"][1] = O.adam_dense(b, s["m"][1], s["v"][1], db, lr, t)
            b0, s["m"][2], s["v"][2] = O.adam_dense(b0, s["m"][2], s["v"][2], db0, lr, t)
        elif opt == "MomentumOptimizer":
            V, s["m"][0] = O.momentum_rows(V, s["m"][0], dV, rows, lr)
            b, s["m"][1] = O.momentum_rows(b, s["m"][1], db, rows, lr)
            b0, s["m"][2] = O.momentum_dense(b0, s["m"][2], db0, lr)
        elif opt == "AdagradOptimizer":
            V, s["m"][0] = O.adagrad_rows(V, s["m"][0], dV, rows, lr)
            b, s["m"][1] = O.adagrad_rows(b, s["m"][1], db, rows, lr)
            b0, s["m"][2] = O.adagrad_dense(b0, s["m"][2], db0, lr)
        else:
            V, b, b0 = O.sgd_dense(V, dV, lr), O.sgd_dense(b, db, lr), O.sgd_dense(b0, db0, lr)
        loss = model.partial_fit({"X": X, "Y": Y})
        assert_close(loss, loss_ref, what="%s loss" % opt)
    got = model.get_weights()
    assert_close(got["feature_embeddings"], V, what="%s V" % opt)
    assert_close(got["feature_bias"].reshape(-1), b, what="%s bias" % opt)
    assert_close(got["bias"], b0, what="%s b0" % opt)
    assert (got["feature_embeddings"][200:] == V_init[200:]).all(), "rows that never appear in a batch must not move"


@pytest.mark.parametrize("dataset,fc,ft,K", [("frappe", 8, 0, 64), ("resturant", 5, 5, 128), ("jiaju", 5, 3, 128)])
def test_hhfm_n_steps_match_oracle(cuda, dataset, fc, ft, K):
    from hhfm_b200.models import OUR
    rng = np.random.default_rng(2)
    n_user, n_item, M = 300, 500, 1000
    lam, lr = 0.01, 0.1
    model = OUR(fc, ft, M, n_user, n_item, K, lr, lam, 'AdagradOptimizer', True, ft > 0)
    V = model.get_weights()["feature_embeddings"].copy(); acc = np.full_like(V, 0.1)
    for step in range(5):
        B = 5000 if step < 4 else 777
        X = np.stack([rng.integers(0, n_user, B), n_user + rng.integers(0, n_item, B)], axis=1)
        F1 = rng.integers(n_user + n_item, M, (B, fc)); F2 = n_user + rng.integers(0, n_item, (B, ft)) if ft else None
        Y = n_user + rng.integers(0, n_item, (B, 10))
        loss_ref, _, _, dV = O.pairrank_loss_grads(V, X, Y, F1, F2, (0, 0, 0), lam)
        V, acc = O.adagrad_dense(V, acc, dV, lr)
        d = {"X": X, "F1": F1, "Y": Y}
        if ft:
            d["F2"] = F2
        loss = model.partial_fit(d)
        assert_close(loss, loss_ref, what="hhfm loss step %d" % step)
    assert_close(model.get_weights()["feature_embeddings"], V, what="hhfm V")
    # PositiveFeadback through the sess shim (OurModel7.py:454) and top-N
    feed = {model.Pos: X[:600], model.Fea: F1[:600]}
    if ft:
        feed[model.Tim] = F2[:600]
    pf = model.sess.run(model.PositiveFeadback, feed_dict=feed)
    ref_pos, _, _, _ = O.pairrank_scores(V, X[:600], None, F1[:600], F2[:600] if ft else None, (0, 0, 0))
    assert pf.shape == (600, 1)
    assert_close(pf[:, 0], ref_pos, what="PositiveFeadback")
    A = np.concatenate([X[:300], F1[:300]] + ([F2[:300]] if ft else []), axis=1)
    ids = model.topk(A, 20)
    Vd = model.get_weights()["feature_embeddings"]
    assert (ids == O.topk_lowest_index(O.hhfm_topk_scores(A, Vd, n_user, n_item, fc, ft), 20)).all()


def test_bpr_and_mf_steps_match_oracle(cuda):
    from hhfm_b200.models import BPR, MF
    rng = np.random.default_rng(3)
    n_user, n_item = 6522, 580
    M, K = n_user + n_item + 100, 128
    bpr = BPR(M, n_user, n_item, K, 0.01, 0.1, 'AdagradOptimizer')
    V = bpr.get_weights()["feature_embeddings"].copy(); acc = np.full_like(V, 1e-8)
    for _ in range(3):
        X = np.stack([rng.integers(0, n_user, 5000), n_user + rng.integers(0, n_item, 5000)], axis=1)
        Y = n_user + rng.integers(0, n_item, (5000, 10))
        loss_ref, _, _, dV = O.pairrank_loss_grads(V, X, Y, None, None, (0, 0, 0), 0.1)
        V, acc = O.adagrad_dense(V, acc, dV, 0.01)
        assert_close(bpr.partial_fit({"X": X, "Y": Y}), loss_ref, what="bpr loss")
    assert_close(bpr.get_weights()["feature_embeddings"], V, what="bpr V")
    A = X[:100]
    ids = bpr.topk(A, 20)
    Vd = bpr.get_weights()["feature_embeddings"]
    assert (ids == O.topk_lowest_index(O.dot_topk_scores(Vd[A[:, 0]], Vd, n_user, n_item), 20)).all()

    mf = MF(M, n_user, n_item, 64, 0.01, 0.01, 1.0, 'AdagradOptimizer', 0, 0)
    V = mf.get_weights()["feature_embeddings"].copy(); acc = np.full_like(V, 1e-8)
    for _ in range(3):
        X = np.stack([rng.integers(0, n_user, 4096), n_user + rng.integers(0, n_item, 4096)], axis=1)
        Y = rng.choice([1.0, -1.0], (4096, 1)).astype(np.float32)
        loss_ref, _, dV = O.mf_loss_grads(X, Y, V, 0.01)
        V, acc = O.adagrad_dense(V, acc, dV, 0.01)
        assert_close(mf.partial_fit({"X": X, "Y": Y}), loss_ref, what="mf loss")
    assert_close(mf.get_weights()["feature_embeddings"], V, what="mf V")
    assert mf.topk(X[:10]).shape == (10, 100)


def test_autograd_functions_match_torch_reference(cuda):
    import torch
    from hhfm_b200 import functional as Fn
    rng = np.random.default_rng(4)
    M, K, B, F = 200, 64, 300, 10
    V0 = rng.normal(0, 0.1, (M, K)).astype(np.float32); b0 = rng.normal(0, 0.1, (M, 1)).astype(np.float32)
    X = rng.integers(0, M, (B, F))
    V = torch.tensor(V0, device=cuda, requires_grad=True); b = torch.tensor(b0, device=cuda, requires_grad=True)
    c = torch.tensor(0.5, device=cuda, requires_grad=True)
    out = Fn.fm_interaction(torch.tensor(X, dtype=torch.int32, device=cuda), V, b, c)
    (out.square().sum()).backward()
    from oracle import torch_cpu as T
    Vr = torch.tensor(V0, requires_grad=True); br = torch.tensor(b0, requires_grad=True); cr = torch.tensor(0.5, requires_grad=True)
    T.fm_out(torch.tensor(X), Vr, br, cr).square().sum().backward()
    assert_close(V.grad.cpu().numpy(), Vr.grad.numpy(), what="autograd gV")
    assert_close(b.grad.cpu().numpy(), br.grad.numpy(), what="autograd gb")
    assert_close(c.grad.item(), cr.grad.item(), what="autograd gb0")
