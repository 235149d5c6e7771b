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


def adagrad_state(model, name="feature_embeddings"):
    return model._opt.state[name][0].detach().cpu().numpy().copy()


def test_fm_steps_match_oracle(cuda):
    """FM.partial_fit (FM.py:168-171) step by step on the model's own trajectory: before each step the oracle is
    given the model's current weights and Adagrad accumulators, then both take the step on the same batch."""
    from conftest import assert_update_close
    from hhfm_b200.models import FM
    rng = np.random.default_rng(0)
    X0, M = frappe_like(rng, 5000)
    K, lam, lr = 64, 0.1, 0.1
    model = FM(10, M, 957, 4082, K, lr, lam, 1, 'AdagradOptimizer', 0, 0)
    model.load_weights({"feature_bias": rng.normal(0, 0.01, (M, 1)).astype(np.float32)})
    accV = np.full((M, K), 0.1, np.float32); accb = np.full(M, 0.1, np.float32); accb0 = np.float32(0.1)
    for step in range(6):
        w = model.get_weights()
        V, b, b0 = w["feature_embeddings"], w["feature_bias"], np.float32(w["bias"])
        if step > 0:
            accV = adagrad_state(model); accb = adagrad_state(model, "feature_bias").reshape(-1); accb0 = np.float32(adagrad_state(model, "bias")[0])
        X, _ = frappe_like(rng, 5000 if step < 5 else 1234)
        Y = rng.choice([1.0, 0.0], (len(X), 1)).astype(np.float32)
        loss_ref, _, dV, db, db0, _ = O.fm_loss_grads(X, Y, V, b, b0, lam)
        V1, _ = O.adagrad_dense(V, accV, dV, lr)
        b1, _ = O.adagrad_dense(b.reshape(-1), accb, db, lr)
        b01, _ = O.adagrad_dense(b0, accb0, db0, lr)
        loss = model.partial_fit({"X": X, "Y": Y})
        assert_close(loss, loss_ref, what="loss step %d" % step)
        got = model.get_weights()
        assert_update_close(got["feature_embeddings"], V1, V, dV, accV, lr, what="V step %d" % step)
        assert_update_close(got["feature_bias"].reshape(-1), b1, b.reshape(-1), db, accb, lr, what="bias step %d" % step)
        assert_update_close(got["bias"], b01, b0, db0, accb0, lr, what="b0 step %d" % step)
    w = model.get_weights()
    V, b, b0 = w["feature_embeddings"], w["feature_bias"], np.float32(w["bias"])
    # forward-only entry used by evaluate_AUC: sess.run(model.out, feed_dict)  (FM.py:313-319)
    out = model.sess.run(model.out, feed_dict={model.train_features: X0[:600], model.train_labels: [[1]] * 600,
                                               model.dropout_keep: 1.0, model.train_phase: False})
    assert out.shape == (600, 1)
    assert_close(out[:, 0], O.fm_forward(X0[:600], V, b, b0)[0], what="predict")
    # top-N through the API, bit-exact lists
    A = X0[:300]
    ids = model.topk(A, 20)
    assert ids.dtype == np.int32 and ids.shape == (300, 20)
    assert (ids == O.topk_lowest_index(O.fm_topk_scores(A, V, b, 957, 4082), 20)).all()


def test_fm_free_running_trajectory_stays_close(cuda):
    """10 un-synchronised steps: per-step losses within 1e-5, final weights within 1e-3 of the total displacement
    (fp32 summation-order differences are amplified by the training dynamics, not by a kernel)."""
    from hhfm_b200.models import FM
    rng = np.random.default_rng(5)
    _, M = frappe_like(rng, 10)
    K, lam, lr = 64, 0.1, 0.1
    model = FM(10, M, 957, 4082, K, lr, lam, 1, 'AdagradOptimizer', 0, 0)
    V = model.get_weights()["feature_embeddings"].copy(); V_init = V.copy()
    b = np.zeros((M, 1), np.float32); b0 = np.float32(0)
    accV = np.full_like(V, 0.1); accb = np.full(M, 0.1, np.float32); accb0 = np.float32(0.1)
    for step in range(10):
        X, _ = frappe_like(rng, 5000)
        Y = rng.choice([1.0, 0.0], (len(X), 1)).astype(np.float32)
        loss_ref, _, dV, db, db0, _ = O.fm_loss_grads(X, Y, V, b, b0, lam)
        V, accV = O.adagrad_dense(V, accV, dV, lr)
        bb, accb = O.adagrad_dense(b.reshape(-1), accb, db, lr); b = bb.reshape(-1, 1)
        b0, accb0 = O.adagrad_dense(b0, accb0, db0, lr)
        assert_close(model.partial_fit({"X": X, "Y": Y}), loss_ref, rtol=2e-5, what="free-running loss step %d" % step)
    got = model.get_weights()["feature_embeddings"]
    assert_close(got - V_init, V - V_init, rtol=1e-3, what="displacement after 10 steps")


@pytest.mark.parametrize("opt", ["AdamOptimizer", "MomentumOptimizer", "GradientDescentOptimizer", "AdagradOptimizer"])
def test_fm_sparse_mode_optimizers(cuda, opt):
    """lamda = 0: IndexedSlices semantics (only touched rows move; TF1 sparse Adam moves every row).  Teacher-forced:
    each step starts from the model's own weights / optimizer slots."""
    from hhfm_b200.models import FM
    rng = np.random.default_rng(1)
    M, K, lr = 300, 32, 0.01
    model = FM(6, M, 50, 100, K, lr, 0.0, 1, opt, 0, 0)
    V_init = model.get_weights()["feature_embeddings"].copy()

    def slots(name, shape, fill):
        st = model._opt.state.get(name)
        if st is None:
            return [np.full(shape, fill, np.float32), np.zeros(shape, np.float32)]
        return [x.detach().cpu().numpy().reshape(shape).copy() if x is not None else None for x in st]

    for t in range(1, 6):
        w = model.get_weights()
        V, b, b0 = w["feature_embeddings"], w["feature_bias"].reshape(-1), np.float32(w["bias"])
        fill = 0.1 if opt == "AdagradOptimizer" else 0.0
        sV, sb, s0 = slots("feature_embeddings", (M, K), fill), slots("feature_bias", (M,), fill), slots("bias", (), fill)
        X = rng.integers(0, 150, (400, 6)); Y = rng.choice([1.0, -1.0], (400, 1)).astype(np.float32)
        loss_ref, _, dV, db, db0, rows = O.fm_loss_grads(X, Y, V, b.reshape(-1, 1), b0, 0.0)
        if opt == "AdamOptimizer":
            V1 = O.adam_dense(V, sV[0], sV[1], dV, lr, t)[0]; b1 = O.adam_dense(b, sb[0], sb[1], db, lr, t)[0]
            b01 = O.adam_dense(b0, s0[0], s0[1], db0, lr, t)[0]
        elif opt == "MomentumOptimizer":
            V1 = O.momentum_rows(V, sV[0], dV, rows, lr)[0]; b1 = O.momentum_rows(b, sb[0], db, rows, lr)[0]
            b01 = O.momentum_dense(b0, s0[0], db0, lr)[0]
        elif opt == "AdagradOptimizer":
            V1 = O.adagrad_rows(V, sV[0], dV, rows, lr)[0]; b1 = O.adagrad_rows(b, sb[0], db, rows, lr)[0]
            b01 = O.adagrad_dense(b0, s0[0], db0, lr)[0]
        else:
            V1, b1, b01 = O.sgd_dense(V, dV, lr), O.sgd_dense(b, db, lr), O.sgd_dense(b0, db0, lr)
        loss = model.partial_fit({"X": X, "Y": Y})
        assert_close(loss, loss_ref, what="%s loss" % opt)
        got = model.get_weights()
        assert_close(got["feature_embeddings"] - V, V1 - V, rtol=2e-5, what="%s V update" % opt)
        assert_close(got["feature_bias"].reshape(-1) - b, b1 - b, rtol=2e-5, what="%s bias update" % opt)
        assert_close(float(got["bias"]) - float(b0), float(b01) - float(b0), rtol=2e-5, what="%s b0 update" % opt)
    assert (model.get_weights()["feature_embeddings"][200:] == V_init[200:]).all(), "rows that never appear must not move"


@pytest.mark.parametrize("dataset,fc,ft,K", [("frappe", 8, 0, 64), ("resturant", 5, 5, 128), ("jiaju", 5, 3, 128)])
def test_hhfm_steps_match_oracle(cuda, dataset, fc, ft, K):
    from conftest import assert_update_close
    from hhfm_b200.models import OUR
    rng = np.random.default_rng(2)
    n_user, n_item, M = 300, 500, 1000
    lam, lr = 0.01, 0.1
    model = OUR(fc, ft, M, n_user, n_item, K, lr, lam, 'AdagradOptimizer', True, ft > 0)
    acc = np.full((M, K), 0.1, np.float32)
    for step in range(5):
        V = model.get_weights()["feature_embeddings"]
        if step > 0:
            acc = adagrad_state(model)
        B = 5000 if step < 4 else 777
        X = np.stack([rng.integers(0, n_user, B), n_user + rng.integers(0, n_item, B)], axis=1)
        F1 = rng.integers(n_user + n_item, M, (B, fc)); F2 = n_user + rng.integers(0, n_item, (B, ft)) if ft else None
        Y = n_user + rng.integers(0, n_item, (B, 10))
        loss_ref, _, _, dV = O.pairrank_loss_grads(V, X, Y, F1, F2, (0, 0, 0), lam)
        V1, _ = O.adagrad_dense(V, acc, dV, lr)
        d = {"X": X, "F1": F1, "Y": Y}
        if ft:
            d["F2"] = F2
        loss = model.partial_fit(d)
        assert_close(loss, loss_ref, what="hhfm loss step %d" % step)
        assert_update_close(model.get_weights()["feature_embeddings"], V1, V, dV, acc, lr, what="hhfm V step %d" % step)
    V = model.get_weights()["feature_embeddings"]
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
    assert (ids == O.topk_lowest_index(O.hhfm_topk_scores(A, V, n_user, n_item, fc, ft), 20)).all()


def test_bpr_and_mf_steps_match_oracle(cuda):
    from conftest import assert_update_close
    from hhfm_b200.models import BPR, MF
    rng = np.random.default_rng(3)
    n_user, n_item = 6522, 580
    M, K = n_user + n_item + 100, 128
    bpr = BPR(M, n_user, n_item, K, 0.01, 0.1, 'AdagradOptimizer')
    acc = np.full((M, K), 1e-8, np.float32)
    for step in range(3):
        V = bpr.get_weights()["feature_embeddings"]
        if step > 0:
            acc = adagrad_state(bpr)
        X = np.stack([rng.integers(0, n_user, 5000), n_user + rng.integers(0, n_item, 5000)], axis=1)
        Y = n_user + rng.integers(0, n_item, (5000, 10))
        loss_ref, _, _, dV = O.pairrank_loss_grads(V, X, Y, None, None, (0, 0, 0), 0.1)
        V1, _ = O.adagrad_dense(V, acc, dV, 0.01)
        assert_close(bpr.partial_fit({"X": X, "Y": Y}), loss_ref, what="bpr loss")
        assert_update_close(bpr.get_weights()["feature_embeddings"], V1, V, dV, acc, 0.01, what="bpr V step %d" % step)
    A = X[:100]
    ids = bpr.topk(A, 20)
    Vd = bpr.get_weights()["feature_embeddings"]
    assert (ids == O.topk_lowest_index(O.dot_topk_scores(Vd[A[:, 0]], Vd, n_user, n_item), 20)).all()
    pf = bpr.sess.run(bpr.PositiveFeadback, feed_dict={bpr.Pos: X[:600]})
    assert_close(pf[:, 0], O.pairrank_scores(Vd, X[:600], None)[0], what="bpr PositiveFeadback")

    mf = MF(M, n_user, n_item, 64, 0.01, 0.01, 1.0, 'AdagradOptimizer', 0, 0)
    acc = np.full((M, 64), 1e-8, np.float32)
    for step in range(3):
        V = mf.get_weights()["feature_embeddings"]
        if step > 0:
            acc = adagrad_state(mf)
        X = np.stack([rng.integers(0, n_user, 4096), n_user + rng.integers(0, n_item, 4096)], axis=1)
        Y = rng.choice([1.0, -1.0], (4096, 1)).astype(np.float32)
        loss_ref, _, dV = O.mf_loss_grads(X, Y, V, 0.01)
        V1, _ = O.adagrad_dense(V, acc, dV, 0.01)
        assert_close(mf.partial_fit({"X": X, "Y": Y}), loss_ref, what="mf loss")
        assert_update_close(mf.get_weights()["feature_embeddings"], V1, V, dV, acc, 0.01, what="mf V step %d" % step)
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


def test_afm_steps_and_topk_match_oracle(cuda, monkeypatch):
    """AFM.partial_fit (AFM.py:205-208) teacher-forced against the oracle + TF1 Adagrad; AFM.topk rank-equivalence."""
    from conftest import assert_update_close
    from hhfm_b200.models import AFM
    rng = np.random.default_rng(7)
    n_user, n_item = 60, 200
    X0, M = frappe_like(rng, 10, n_user=n_user, n_item=n_item, ctx=(7, 2, 3, 9))
    K, lr, lam = 64, 0.1, 100.0
    F = X0.shape[1]
    model = AFM(n_user, n_item, M, 1, [K, K], 'relu', lr, lam, [1, 1], 'AdagradOptimizer', 0.999, F)
    names = ["feature_embeddings", "feature_bias", "bias", "attention_W", "attention_b", "attention_p", "prediction"]
    for step in range(4):
        w = model.get_weights()
        acc = {k: (model._opt.state[k][0].detach().cpu().numpy().reshape(np.asarray(w[k]).shape).copy() if k in model._opt.state
                   else np.full(np.asarray(w[k]).shape, 0.1, np.float32)) for k in names}
        X, _ = frappe_like(rng, 3000 if step < 3 else 517, n_user=n_user, n_item=n_item, ctx=(7, 2, 3, 9))
        Y = rng.choice([1.0, -1.0], (len(X), 1)).astype(np.float32)
        loss_ref, _, g = O.afm_loss_grads(X, Y, w, lam)
        loss = model.partial_fit({"X": X, "Y": Y})
        assert_close(loss, loss_ref, what="afm loss step %d" % step)
        got = model.get_weights()
        for k in names:
            gk = np.asarray(g[k], np.float32).reshape(np.asarray(w[k]).shape)
            w1, _ = O.adagrad_dense(np.asarray(w[k], np.float32), acc[k], gk, lr)
            if k in ("feature_embeddings", "feature_bias"):
                assert_update_close(got[k], w1, w[k], gk, acc[k], lr, rtol=3e-5, what="afm %s step %d" % (k, step))
            else:
                # Small dense variables: their gradients are sums over B*P terms that largely cancel (softmax gradients
                # sum to zero over the pairs), so single elements are fp32 rounding noise in the oracle and on the device
                # alike; the update is compared norm-wise (1e-4 of the largest update of the tensor, plus a 1e-8 floor: at
                # the reference init Z ~ b for every pair, so d attention_b is an exactly cancelling sum).
                d_ref = np.asarray(w1, np.float64) - np.asarray(w[k], np.float64)
                d_got = np.asarray(got[k], np.float64).reshape(d_ref.shape) - np.asarray(w[k], np.float64)
                ulp = 4 * 1.2e-7 * float(np.abs(np.asarray(w[k], np.float64)).max())     # 4 ulp of the tensor's largest weight
                assert np.abs(d_got - d_ref).max() <= 1e-4 * np.abs(d_ref).max() + 1e-8 + ulp, (k, step, np.abs(d_got - d_ref).max(), np.abs(d_ref).max())
    # predict through the sess shim and top-N (rank-equivalent to the reference formula)
    w = model.get_weights()
    out = model.sess.run(model.out, feed_dict={model.train_features: X[:300], model.train_labels: [[1]] * 300,
                                               model.dropout_keep: [1.0, 1.0], model.train_phase: False})
    assert_close(out[:, 0], O.afm_forward(X[:300], w)[0], what="afm predict")
    A = X[:40]
    ids = model.topk(A, 20)                                   # item-separable scorer (csrc/afm_topn.cu)
    monkeypatch.setenv("HHFM_AFM_TOPN_SEPARABLE", "0")
    ids_full = model.topk(A, 20)                              # every (row, item) pair through the forward kernel
    monkeypatch.delenv("HHFM_AFM_TOPN_SEPARABLE")
    assert (ids == ids_full).mean() > 0.99
    ref = O.afm_topk_scores(A, w, n_user, n_item)
    want = O.topk_lowest_index(ref, 20)
    for r in range(len(A)):
        if (ids[r] == want[r]).all():
            continue
        # lists may differ only between items whose reference scores are numerically indistinguishable
        kth = ref[r, want[r, -1]]
        tol = 2e-5 * max(abs(kth), float(np.sqrt(np.mean(ref[r] ** 2))))
        for a_, b_ in zip(ids[r], want[r]):
            assert abs(ref[r, a_] - ref[r, b_]) <= tol, (r, a_, b_, ref[r, a_], ref[r, b_])


def test_dfm_steps_and_topk_match_oracle(cuda, monkeypatch):
    """DeepFM.partial_fit (DFM.py:216-219) teacher-forced against the oracle + TF1 Adagrad; DeepFM.topk (DFM.py:220-231)."""
    from conftest import assert_update_close
    from hhfm_b200.models import DeepFM
    rng = np.random.default_rng(11)
    n_user, n_item = 60, 200
    X0, M = frappe_like(rng, 10, n_user=n_user, n_item=n_item, ctx=(7, 2, 3, 9))
    K, lr, lam = 64, 0.01, 0.01
    F = X0.shape[1]
    layers = [150, 200, 150]
    model = DeepFM(n_user, n_item, M, F, K, layers, 'relu', lr, 0, lam)
    names = list(model.weights.keys())
    acc_name = {"feature_embeddings": "feature_embeddings", "feature_bias": "feature_bias"}
    for step in range(4):
        w = model.get_weights()
        acc_reg = model._opt.state["dense_reg"][0].cpu().numpy().copy() if "dense_reg" in model._opt.state else None
        acc_bias = model._opt.state["dense_bias"][0].cpu().numpy().copy() if "dense_bias" in model._opt.state else None
        flat_off = {}
        off = 0
        for i in range(3):
            flat_off["layer_%d" % i] = ("r", off); off += w["layer_%d" % i].size
        flat_off["concat_projection"] = ("r", off)
        off = 0
        for i in range(3):
            flat_off["bias_%d" % i] = ("b", off); off += w["bias_%d" % i].size
        flat_off["concat_bias"] = ("b", off)
        X, _ = frappe_like(rng, 3000 if step < 3 else 517, n_user=n_user, n_item=n_item, ctx=(7, 2, 3, 9))
        Y = rng.choice([1.0, -1.0], (len(X), 1)).astype(np.float32)
        loss_ref, _, g = O.dfm_loss_grads(X, Y, w, lam)
        acc_tab = {k: (model._opt.state[k][0].cpu().numpy().copy() if k in model._opt.state else np.full(w[k].shape, 0.1, np.float32))
                   for k in ("feature_embeddings", "feature_bias")}
        loss = model.partial_fit({"X": X, "Y": Y})
        assert_close(loss, loss_ref, what="dfm loss step %d" % step)
        got = model.get_weights()
        for k in names:
            wk = np.asarray(w[k], np.float32)
            gk = np.asarray(g[k], np.float32).reshape(wk.shape)
            if k in acc_tab:
                acc = acc_tab[k]
            else:
                kind, o_ = flat_off[k]
                src = acc_reg if kind == "r" else acc_bias
                acc = np.full(wk.shape, 0.1, np.float32) if src is None else src[o_:o_ + wk.size].reshape(wk.shape)
            w1, _ = O.adagrad_dense(wk, acc, gk, lr)
            assert_update_close(np.asarray(got[k]).reshape(wk.shape), w1, wk, gk, acc, lr, rtol=3e-5,
                                atol=4 * 1.2e-7 * float(np.abs(wk).max()), what="dfm %s step %d" % (k, step))
    w = model.get_weights()
    out = model.sess.run(model.out, feed_dict={model.feat_index: X[:300], model.label: [[1]] * 300})
    assert_close(out[:, 0], O.dfm_forward(X[:300], w)[0], what="dfm predict")
    A = X[:40]
    rows = np.repeat(A[:, None, :], n_item, axis=1); rows[:, :, 1] = n_user + np.arange(n_item)[None, :]
    ref = O.dfm_forward(rows.reshape(-1, F), w)[0].reshape(len(A), n_item)
    want = O.topk_lowest_index(ref, 20)
    # the item-separable evaluator's scores against the op-by-op forward of every expanded row, on both GEMM paths
    import torch
    from hhfm_b200 import _lib
    from hhfm_b200.engine import cur_stream, ptr
    for tc in ("1", "0"):
        monkeypatch.setenv("HHFM_DFM_TC", tc)
        A_dev, stride = model._topn.upload_rows(A, model._M)
        need = int(_lib.load().hhfm_workspace_bytes_dfm_topn(len(A), n_item, F, K, len(layers), model._sizes.ctypes.data)) // 4
        ws = torch.empty(need, device=cuda); sc = torch.full((len(A), n_item), float("nan"), device=cuda)
        _lib.call("hhfm_dfm_topn_scores", ptr(A_dev), stride, len(A), F, 1, ptr(model.weights["feature_embeddings"]),
                  ptr(model.weights["feature_bias"]), model._M, K, ptr(model._params), len(layers), model._sizes.ctypes.data,
                  n_user, n_item, ptr(ws), ptr(sc), cur_stream())
        assert_close(sc.cpu().numpy(), ref, rtol=2e-5, what="dfm separable scores (HHFM_DFM_TC=%s)" % tc)
    monkeypatch.delenv("HHFM_DFM_TC")
    for sep in ("1", "0"):          # separable evaluator / every expanded row through hhfm_dfm_fwd
        monkeypatch.setenv("HHFM_DFM_TOPN_SEPARABLE", sep)
        ids = model.topk(A, 20)
        for r in range(len(A)):
            if (ids[r] == want[r]).all():
                continue
            tol = 2e-5 * max(abs(ref[r, want[r, -1]]), float(np.sqrt(np.mean(ref[r] ** 2))))
            for a_, b_ in zip(ids[r], want[r]):
                assert abs(ref[r, a_] - ref[r, b_]) <= tol, (r, a_, b_, ref[r, a_], ref[r, b_])


def test_wd_steps_and_topk_match_oracle(cuda):
    """Wide&Deep (WDMF.py:51-126 restated; parity unpinned against TF, pinned against the oracle): teacher-forced full-batch
    steps -- mean log-loss, Adagrad on the DNN half / embeddings, FTRL on the wide half -- then predict_proba and topk."""
    from conftest import assert_update_close
    from hhfm_b200.models import WD
    rng = np.random.default_rng(13)
    n_user, n_item, F, K = 40, 120, 5, 16
    M = 300
    hidden = [48, 32, 16]
    model = WD(F, n_user, n_item, features_M=M, hidden_units=hidden, embedding_dim=K, cross_buckets=97, steps=1)
    # the estimator starts the wide half and the biases at zero; give them values so that every term of the graph is exercised
    init = model.get_weights()
    for k in ("wide_linear", "wide_cross", "wide_bias", "bias_0", "bias_1", "bias_2", "logits_b"):
        init[k] = rng.normal(0, 0.1, np.asarray(init[k]).shape).astype(np.float32)
    model.load_weights(init)
    adagrad = ["feature_embeddings", "layer_0", "layer_1", "layer_2", "bias_0", "bias_1", "bias_2", "logits_w", "logits_b"]
    ftrl = ["wide_linear", "wide_cross", "wide_bias"]
    acc = {k: np.full(np.asarray(init[k]).shape, 0.1, np.float32) for k in adagrad + ftrl}
    lin = {k: np.zeros(np.asarray(init[k]).shape, np.float32) for k in ftrl}
    lr_w = min(0.2, 1.0 / np.sqrt(F + F * (F - 1) // 2))
    assert abs(model.linear_learning_rate - lr_w) < 1e-12 and model.dnn_learning_rate == 0.05
    for step in range(3):
        B = 2500 if step < 2 else 333
        X = np.stack([rng.integers(0, n_user, B), n_user + rng.integers(0, n_item, B), 160 + rng.integers(0, 7, B),
                      170 + rng.integers(0, 2, B), 180 + rng.integers(0, 100, B)], axis=1)
        Y = rng.choice([1.0, 0.0], B).astype(np.float32)
        w = model.get_weights()
        loss_ref, z_ref, g = O.wd_loss_grads(X, Y, w)
        loss = model.partial_fit(X, Y)
        assert_close(loss, loss_ref, what="wd loss step %d" % step)
        got = model.get_weights()
        for k in adagrad:
            wk = np.asarray(w[k], np.float32); gk = np.asarray(g[k], np.float32).reshape(wk.shape)
            w1, a1 = O.adagrad_dense(wk, acc[k], gk, 0.05)
            assert_update_close(np.asarray(got[k]).reshape(wk.shape), w1, wk, gk, acc[k], 0.05, rtol=3e-5,
                                atol=4 * 1.2e-7 * float(np.abs(wk).max()), what="wd %s step %d" % (k, step))
            acc[k] = a1
        for k in ftrl:
            wk = np.asarray(w[k], np.float32); gk = np.asarray(g[k], np.float32).reshape(wk.shape)
            w1, a1, z1 = O.ftrl_dense(wk, acc[k], lin[k], gk, lr_w)
            d_ref = w1.astype(np.float64) - wk
            d_got = np.asarray(got[k], np.float64).reshape(wk.shape) - wk
            scale = max(float(np.abs(d_ref).max()), 1e-12)
            assert np.abs(d_got - d_ref).max() <= 3e-5 * scale + 4 * 1.2e-7 * float(np.abs(wk).max()), (k, step, np.abs(d_got - d_ref).max(), scale)
            acc[k] = a1; lin[k] = z1
    w = model.get_weights()
    z = O.wd_forward(X[:200], w)[0]
    proba = model.predict(X[:200])
    assert proba.shape == (200, 2)
    assert_close(proba[:, 1], 1.0 / (1.0 + np.exp(-z.astype(np.float64))), what="wd predict_proba")
    assert np.allclose(proba.sum(axis=1), 1.0, atol=1e-6)
    A = X[:30]
    ids = model.topk(A, 20)
    rows = np.repeat(A[:, None, :], n_item, axis=1); rows[:, :, 1] = n_user + np.arange(n_item)[None, :]
    ref = O.wd_forward(rows.reshape(-1, F), w)[0].reshape(len(A), n_item)
    want = O.topk_lowest_index(ref, 20)
    for r in range(len(A)):
        if (ids[r] == want[r]).all():
            continue
        tol = 2e-5 * max(abs(ref[r, want[r, -1]]), float(np.sqrt(np.mean(ref[r] ** 2))))
        for a_, b_ in zip(ids[r], want[r]):
            assert abs(ref[r, a_] - ref[r, b_]) <= tol, (r, a_, b_, ref[r, a_], ref[r, b_])


def _cars2_weights(rng, n_ui, M, D):
    Dc, Dp, Dq = int(D / 2.5), int(D / 5), int(D / 2.5)
    return dict(UI=rng.normal(0, 0.1, (n_ui, D)).astype(np.float32), Context=rng.normal(0, 0.1, (M, Dc)).astype(np.float32),
                W=rng.normal(0, 0.1, (D, Dp, Dc)).astype(np.float32), Z=rng.normal(0, 0.1, (D, Dq, Dc)).astype(np.float32),
                A=rng.normal(0, 0.3, Dp).astype(np.float32), B=rng.normal(0, 0.3, Dq).astype(np.float32))


@pytest.mark.parametrize("D,NG", [(128, 1), (64, 3), (20, 2)])
def test_cars2_steps_and_topk_match_oracle(cuda, D, NG):
    """CARS2.partial_fit (CARS2.py:168-171) teacher-forced against the oracle + TF1 Adagrad, PositiveFeadback through the
    sess shim and CARS2.topk (CARS2.py:171-187)."""
    from conftest import assert_update_close
    from hhfm_b200.models import CARS2
    rng = np.random.default_rng(D + NG)
    n_user, n_item, M, lr, lam = 40, 90, 25, 0.05, 0.001
    model = CARS2(M, n_user, n_item, D, lr, lam, 'AdagradOptimizer')
    model.load_weights(_cars2_weights(rng, n_user + n_item, M, D))
    names = ["UI", "Context", "W", "Z", "A", "B"]
    for step in range(3):
        w = model.get_weights()
        acc_flat = model._opt.state["params"][0].cpu().numpy().copy() if "params" in model._opt.state else None
        B = 2000 if step < 2 else 317
        Pos = np.stack([rng.integers(0, n_user, B), n_user + rng.integers(0, n_item, B)], axis=1)
        Neg = n_user + rng.integers(0, n_item, (B, NG)); Fea = rng.integers(0, M, B)
        loss_ref, pos_ref, g = O.cars2_loss_grads(Pos, Fea, Neg, w, lam)
        fb = model.sess.run(model.PositiveFeadback, feed_dict={model.Pos: Pos, model.Fea: Fea})
        assert_close(fb[:, 0], pos_ref, rtol=2e-5, what="cars2 PositiveFeadback step %d" % step)
        loss = model.partial_fit({"X": Pos, "F1": Fea, "Y": Neg})
        assert_close(loss, loss_ref, what="cars2 loss step %d" % step)
        got = model.get_weights()
        off = 0
        for k in names:
            wk = np.asarray(w[k], np.float32); gk = np.asarray(g[k], np.float32).reshape(wk.shape)
            acc = np.full(wk.shape, 0.1, np.float32) if acc_flat is None else acc_flat[off:off + wk.size].reshape(wk.shape)
            off += wk.size
            w1, _ = O.adagrad_dense(wk, acc, gk, lr)
            # d B_q contracts Z with dT = Delta^T C (two fp32 stages), the oracle's einsum contracts four factors at once:
            # both are fp32, in a different association
            assert_update_close(got[k], w1, wk, gk, acc, lr, rtol=1e-4 if k in ("A", "B") else 3e-5,
                                atol=4 * 1.2e-7 * float(np.abs(wk).max()), what="cars2 %s step %d" % (k, step))
    w = model.get_weights()
    users = rng.integers(0, n_user, 50); fea = rng.integers(0, M, 50)
    ids = model.topk({"X": users, "F1": fea}, 20)
    ref = O.cars2_topk_scores(users, fea, w, n_user, n_item)
    want = O.topk_lowest_index(ref, 20)
    for r in range(len(users)):
        if (ids[r] == want[r]).all():
            continue
        tol = 2e-5 * max(abs(ref[r, want[r, -1]]), float(np.sqrt(np.mean(ref[r] ** 2))))
        for a_, b_ in zip(ids[r], want[r]):
            assert abs(ref[r, a_] - ref[r, b_]) <= tol, (r, a_, b_, ref[r, a_], ref[r, b_])


def test_models_match_committed_golden_vectors(cuda):
    """DeepFM / Wide&Deep / CARS2 forward passes on the committed fixtures of tests/golden/oracle_vectors.npz (weights and
    inputs injected, outputs frozen when the oracle was pinned)."""
    import os
    from hhfm_b200.models import CARS2, DeepFM, WD
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_vectors.npz"))
    n_user, n_item, M, K, B, F, NG = [int(x) for x in g["dims"]]
    X, V = g["X"], g["V"]
    layers = [12, 10, 6]
    dfm = DeepFM(n_user, n_item, M, F, K, layers, 'relu', 0.01, 0, 0.01)
    w = {"feature_embeddings": V}
    for k in dfm.weights:
        if k != "feature_embeddings":
            w[k] = g["dfm_w_" + k]
    dfm.load_weights(w)
    assert_close(dfm.predict(X)[:, 0], g["dfm_out"], rtol=2e-5, what="golden dfm out")
    wd = WD(F, n_user, n_item, features_M=M, hidden_units=layers, embedding_dim=K, cross_buckets=97, steps=1)
    w = {"feature_embeddings": V, "wide_linear": g["wd_wide_linear"], "wide_cross": g["wd_wide_cross"], "wide_bias": np.float32(0.03),
         "logits_w": g["wd_logits_w"], "logits_b": np.float32(-0.02)}
    for i in range(3):
        w["layer_%d" % i] = g["dfm_w_layer_%d" % i]; w["bias_%d" % i] = g["dfm_w_bias_%d" % i]
    wd.load_weights(w)
    assert_close(wd.score_device(wd._upload_rows(X)).cpu().numpy(), g["wd_logit"], rtol=2e-5, what="golden wd logit")
    assert_close(wd.predict(X)[:, 1], 1.0 / (1.0 + np.exp(-g["wd_logit"].astype(np.float64))), rtol=2e-5, what="golden wd proba")
    D = g["cars2_w_UI"].shape[1]
    c2 = CARS2(g["cars2_w_Context"].shape[0], n_user, n_item, D, 0.05, 0.001, 'AdagradOptimizer')
    c2.load_weights({k: g["cars2_w_" + k] for k in ("UI", "Context", "W", "Z", "A", "B")})
    fb = c2.sess.run(c2.PositiveFeadback, feed_dict={c2.Pos: X[:, :2], c2.Fea: g["cars2_Fea"]})
    assert_close(fb[:, 0], g["cars2_feedback"], rtol=2e-5, what="golden cars2 PositiveFeadback")


def test_cars2_momentum_without_l2_raises_instead_of_moving_untouched_rows(cuda):
    """CARS2.py with lamda == 0 gives UI / Context IndexedSlices gradients, i.e. TF's SparseApplyMomentum; the dense Momentum
    kernel would keep moving untouched rows, so the combination must refuse to run."""
    from hhfm_b200.models import CARS2
    rng = np.random.default_rng(0)
    n_user, n_item, M, D, B = 20, 30, 10, 20, 64
    model = CARS2(M, n_user, n_item, D, 0.05, 0.0, 'MomentumOptimizer')
    Pos = np.stack([rng.integers(0, n_user, B), n_user + rng.integers(0, n_item, B)], axis=1)
    with pytest.raises(NotImplementedError):
        model.partial_fit({"X": Pos, "F1": rng.integers(0, M, B), "Y": n_user + rng.integers(0, n_item, (B, 2))})


def test_invalidate_drops_the_cached_item_operand_after_an_in_place_weight_edit(cuda):
    """The tensor-core top-N caches the bf16 item operand per weight version; `invalidate()` is the hook for edits of
    `model.weights[...]` that bypass the training step."""
    import torch
    from hhfm_b200.models import BPR
    rng = np.random.default_rng(3)
    n_user, n_item, K, C = 64, 4000, 64, 1100
    m = BPR(n_user + n_item, n_user, n_item, K, 0.05, 0.0, 'AdagradOptimizer')
    m.topn_method = "tc"
    A = np.stack([rng.integers(0, n_user, C), n_user + rng.integers(0, n_item, C)], axis=1)
    V = m.get_weights()["feature_embeddings"]
    first = m.topk(A, 10)
    assert (first == O.topk_lowest_index(V[A[:, 0]] @ V[n_user:].T, 10)).mean() > 0.99      # fp32 order of the dot products aside
    with torch.no_grad():
        m.weights["feature_embeddings"][n_user:] *= -1.0
    m.invalidate()
    second = m.topk(A, 10)
    V2 = m.get_weights()["feature_embeddings"]
    assert (second == O.topk_lowest_index(V2[A[:, 0]] @ V2[n_user:].T, 10)).mean() > 0.99
    assert (first != second).mean() > 0.9
