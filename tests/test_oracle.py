"""CPU: the oracle against (a) torch.autograd on the op-for-op torch mirror of the TF graphs, (b) its own committed
golden vectors.  The reference cannot run here (no TensorFlow 1.x), so this is what pins the restatement's algebra."""
import os

import numpy as np
import pytest
import torch

from conftest import assert_close
from oracle import hhfm_oracle as O
from oracle import torch_cpu as T

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_reproduces_its_golden_vectors():
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(HERE, "golden", "make_oracle_goldens.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    gold = np.load(os.path.join(HERE, "golden", "oracle_vectors.npz"))
    now = mk.cases()
    assert sorted(now) == sorted(gold.files)
    for k in gold.files:
        a, b = np.asarray(now[k]), gold[k]
        if a.dtype.kind == "f":
            assert_close(a, b, rtol=2e-6, what=k)      # BLAS/libm differences between hosts only
        else:
            assert (a == b).all(), k


@pytest.fixture
def prob():
    rng = np.random.default_rng(0)
    M, K, B, F, NG = 200, 16, 64, 6, 5
    V = rng.normal(0, 0.1, (M, K)).astype(np.float32); b = rng.normal(0, 0.1, (M, 1)).astype(np.float32)
    X = rng.integers(0, M, (B, F)); X[:, 3] = X[:, 2]
    Y = rng.choice([1.0, -1.0], (B, 1)).astype(np.float32)
    return rng, M, K, B, F, NG, V, b, X, Y


def test_fm_gradients_agree_with_autograd(prob):
    rng, M, K, B, F, NG, V, b, X, Y = prob
    loss, out, dV, db, db0, _ = O.fm_loss_grads(X, Y, V, b, np.float32(0.3), lamda=0.1)
    tV = torch.tensor(V, requires_grad=True); tb = torch.tensor(b, requires_grad=True); tb0 = torch.tensor(0.3, requires_grad=True)
    l, o = T.fm_loss(torch.tensor(X), torch.tensor(Y), tV, tb, tb0, 0.1); l.backward()
    assert_close(loss, l.item(), what="loss"); assert_close(out, o.detach().numpy()[:, 0], what="out")
    assert_close(dV, tV.grad.numpy(), what="dV"); assert_close(db, tb.grad.numpy()[:, 0], what="db")
    assert_close(db0, tb0.grad.item(), what="db0")


@pytest.mark.parametrize("pools", [(0, 0, 0), (1, 1, 1), (2, 2, 2), (1, 0, 2), (0, 2, 1)])
def test_pairrank_gradients_agree_with_autograd(prob, pools):
    rng, M, K, B, F, NG, V, b, X, Y = prob
    Pos = rng.integers(0, M, (B, 2)); Fea = rng.integers(0, M, (B, 3)); Fea[:, 1] = Fea[:, 0]
    Tim = rng.integers(0, M, (B, 2)); Neg = rng.integers(0, M, (B, NG)); Neg[:, 1] = Neg[:, 0]
    loss, pos, neg, dV = O.pairrank_loss_grads(V, Pos, Neg, Fea, Tim, pools, 0.01)
    tV = torch.tensor(V, requires_grad=True)
    l, p, n = T.pairrank_loss(tV, torch.tensor(Pos), torch.tensor(Neg), torch.tensor(Fea), torch.tensor(Tim), pools, 0.01)
    l.backward()
    assert_close(loss, l.item(), what="loss"); assert_close(dV, tV.grad.numpy(), what="dV")
    loss, pos, neg, dV = O.pairrank_loss_grads(V, Pos, Neg, None, None, (0, 0, 0), 0.1)       # BPR
    tV = torch.tensor(V, requires_grad=True)
    l, _, _ = T.pairrank_loss(tV, torch.tensor(Pos), torch.tensor(Neg), None, None, (0, 0, 0), 0.1); l.backward()
    assert_close(dV, tV.grad.numpy(), what="bpr dV")


def test_afm_gradients_agree_with_autograd(prob):
    rng, M, K, B, F, NG, V, b, X, Y = prob
    w = dict(feature_embeddings=V, feature_bias=b, bias=np.float32(0.1), attention_W=rng.normal(0, 0.2, (K, K)).astype(np.float32),
             attention_b=rng.normal(0, 0.2, (1, K)).astype(np.float32), attention_p=rng.normal(0, 1, (K,)).astype(np.float32),
             prediction=rng.normal(1, 0.1, (K, 1)).astype(np.float32))
    loss, out, g = O.afm_loss_grads(X, Y, w, 100.0)
    tw = {k: torch.tensor(v, requires_grad=True) for k, v in w.items()}
    l, o = T.afm_loss(torch.tensor(X), torch.tensor(Y), tw, 100.0); l.backward()
    assert_close(loss, l.item(), what="afm loss"); assert_close(out, o.detach().numpy()[:, 0], what="afm out")
    for k in g:
        assert_close(np.asarray(g[k]).reshape(-1), tw[k].grad.numpy().reshape(-1), rtol=2e-5, what="afm grad " + k)


def test_topk_order_is_descending_with_lowest_index_ties():
    s = np.array([[1.0, 3.0, 3.0, -0.0, 0.0, 2.0]], np.float32)
    assert O.topk_lowest_index(s, 6).tolist() == [[1, 2, 5, 0, 3, 4]]
    rng = np.random.default_rng(1)
    s = rng.integers(-3, 4, (50, 200)).astype(np.float32)
    ts, ti = torch.topk(torch.tensor(s), 200, dim=1, sorted=True)
    got = O.topk_lowest_index(s, 200)
    assert (np.take_along_axis(s, got, 1) == ts.numpy()).all()
    for r in range(50):                                     # equal scores appear in ascending index order
        v = s[r, got[r]]
        for a, c in zip(range(199), range(1, 200)):
            if v[a] == v[c]:
                assert got[r, a] < got[r, c]


def test_tf1_optimizer_restatements():
    w = np.array([1.0, -2.0], np.float32); g = np.array([0.5, 0.0], np.float32)
    w1, a1 = O.adagrad_dense(w, np.full(2, 0.1, np.float32), g, 0.1)
    assert_close(a1, [0.35, 0.1]); assert_close(w1, [1.0 - 0.1 * 0.5 / np.sqrt(0.35), -2.0])
    w1, m1, v1 = O.adam_dense(w, np.zeros(2, np.float32), np.zeros(2, np.float32), g, 0.01, 1)
    assert_close(w1, [1.0 - 0.01 * np.sqrt(1 - 0.999) / (1 - 0.9) * 0.05 / (np.sqrt(0.00025) + 1e-8), -2.0])
    w1, a1 = O.momentum_dense(w, np.array([1.0, 1.0], np.float32), g, 0.1)
    assert_close(a1, [1.45, 0.95]); assert_close(w1, [1.0 - 0.145, -2.0 - 0.095])


def test_dfm_gradients_agree_with_autograd():
    """DeepFM restatement (DFM.py:104-152): every gradient against torch.autograd in float64."""
    rng = np.random.default_rng(0)
    M, K, F, B = 50, 8, 5, 40
    dims = [F * K, 7, 9, 6]
    w = {"feature_embeddings": rng.normal(0, .3, (M, K)).astype(np.float32), "feature_bias": rng.uniform(0, 1, (M, 1)).astype(np.float32),
         "concat_projection": rng.normal(0, .3, (F + K + dims[-1], 1)).astype(np.float32), "concat_bias": np.float32(0.01)}
    for i in range(3):
        w["layer_%d" % i] = rng.normal(0, .3, (dims[i], dims[i + 1])).astype(np.float32)
        w["bias_%d" % i] = rng.normal(0, .3, (1, dims[i + 1])).astype(np.float32)
    X = rng.integers(0, M, (B, F)); Y = rng.choice([1., -1.], (B, 1)).astype(np.float32)
    loss, out, g = O.dfm_loss_grads(X, Y, w, 0.01)
    tw = {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=True) for k, v in w.items()}
    E = tw["feature_embeddings"][torch.tensor(X)]
    y1 = tw["feature_bias"].reshape(-1)[torch.tensor(X)]
    S = E.sum(1); y2 = 0.5 * (S * S - (E * E).sum(1))
    h = E.reshape(B, -1)
    for i in range(3):
        h = torch.relu(h @ tw["layer_%d" % i] + tw["bias_%d" % i])
    o = (torch.cat([y1, y2, h], 1) @ tw["concat_projection"]).reshape(-1) + tw["concat_bias"]
    L = 0.5 * ((torch.tensor(Y.reshape(-1), dtype=torch.float64) - o) ** 2).sum() + 0.005 * (
        (tw["concat_projection"] ** 2).sum() + sum((tw["layer_%d" % i] ** 2).sum() for i in range(3)))
    L.backward()
    assert_close(loss, L.item(), what="dfm loss"); assert_close(out, o.detach().numpy(), what="dfm out")
    for k in g:
        assert_close(np.asarray(g[k], np.float64).reshape(-1), tw[k].grad.numpy().reshape(-1), rtol=2e-5, what="dfm grad " + k)


def test_cars2_gradients_agree_with_autograd():
    """CARS2 restatement (CARS2.py:85-123): every gradient against torch.autograd in float64."""
    rng = np.random.default_rng(1)
    nu, ni, M, D = 20, 30, 12, 10
    Dc, Dp, Dq = 4, 2, 4
    B, NG = 50, 2
    w = dict(UI=rng.normal(0, .3, (nu + ni, D)).astype(np.float32), Context=rng.normal(0, .3, (M, Dc)).astype(np.float32),
             W=rng.normal(0, .3, (D, Dp, Dc)).astype(np.float32), Z=rng.normal(0, .3, (D, Dq, Dc)).astype(np.float32),
             A=rng.normal(0, .3, Dp).astype(np.float32), B=rng.normal(0, .3, Dq).astype(np.float32))
    Pos = np.stack([rng.integers(0, nu, B), nu + rng.integers(0, ni, B)], 1)
    Neg = nu + rng.integers(0, ni, (B, NG)); Fea = rng.integers(0, M, B)
    loss, pos, g = O.cars2_loss_grads(Pos, Fea, Neg, w, 0.01)
    tw = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in w.items()}

    def fb(u, it, C):
        pik = torch.einsum("bd,dpc,bc->bp", u, tw["W"], C); q = torch.einsum("bd,dqc,bc->bq", it, tw["Z"], C)
        return (u * it).sum(1) + (pik * tw["A"]).sum(1) + (q * tw["B"]).sum(1)
    u = tw["UI"][torch.tensor(Pos[:, 0])]; ip = tw["UI"][torch.tensor(Pos[:, 1])]
    ineg = tw["UI"][torch.tensor(Neg)].sum(1); C = tw["Context"][torch.tensor(Fea)]
    L = -torch.log(torch.sigmoid(fb(u, ip, C) - fb(u, ineg, C))).sum() + 0.005 * sum((v ** 2).sum() for v in tw.values())
    L.backward()
    assert_close(loss, L.item(), what="cars2 loss"); assert_close(pos, fb(u, ip, C).detach().numpy(), what="cars2 feedback")
    for k in g:
        assert_close(np.asarray(g[k], np.float64).reshape(-1), tw[k].grad.numpy().reshape(-1), rtol=2e-5, what="cars2 grad " + k)
    # the ranking used by CARS2.topk is the dot product with u + T c up to a per-row constant
    users = rng.integers(0, nu, 7); fea = rng.integers(0, M, 7)
    ref = O.cars2_topk_scores(users, fea, w, nu, ni)
    T_ = np.einsum("q,dqc->dc", w["B"], w["Z"])
    q = w["UI"][users] + w["Context"][fea] @ T_.T
    alt = q @ w["UI"][nu:nu + ni].T
    d = ref - alt
    assert np.abs(d - d[:, :1]).max() < 1e-5                  # differs by a constant per row only


def test_hashed_negative_sampler_restatement():
    """The counter-based sampler (device: csrc/sampler.cu): range, rejection, determinism, order independence."""
    rng = np.random.default_rng(3)
    n_user, n_item, span = 10, 25, 35
    codes = np.unique(np.array([k * span + n_user + i for k in range(6) for i in rng.choice(n_item, 20, replace=False)], np.int64))
    key_id = rng.integers(-1, 6, 400).astype(np.int64)
    a = O.sample_negative_hashed(key_id, 5, n_user, n_item, codes, span, 1234)
    b = O.sample_negative_hashed(key_id, 5, n_user, n_item, codes, span, 1234)
    assert (a == b).all() and a.min() >= n_user and a.max() < n_user + n_item
    cs = set(codes.tolist())
    assert not any((int(k) * span + int(it)) in cs for k, row in zip(key_id, a) if k >= 0 for it in row)
    c = O.sample_negative_hashed(key_id, 5, n_user, n_item, codes, span, 1235)
    assert (a != c).mean() > 0.5                                # another seed, another sample
    # a draw depends on (seed, cell, attempt) only: sampling a prefix of the rows gives the same cells
    d = O.sample_negative_hashed(key_id[:100], 5, n_user, n_item, codes, span, 1234)
    assert (d == a[:100]).all()


def test_wd_gradients_match_autograd_and_ftrl_reduces_to_adagrad():
    """Wide&Deep restatement (WDMF.py:51-126; parity unpinned): analytic gradients against torch.autograd in float64, the
    cross-bucket hash against a scalar splitmix64, and FTRL's first step from (n, z, w) = (0.1, 0, 0) against Adagrad."""
    import torch
    rng = np.random.default_rng(0)
    M, K, F, B, NB = 60, 8, 4, 50, 97
    layers = [12, 10, 6]
    w = dict(feature_embeddings=rng.normal(0, 0.3, (M, K)).astype(np.float32), wide_linear=rng.normal(0, 0.3, (M,)).astype(np.float32),
             wide_cross=rng.normal(0, 0.3, (6, NB)).astype(np.float32), wide_bias=np.float32(0.1),
             logits_w=rng.normal(0, 0.3, (layers[-1], 1)).astype(np.float32), logits_b=np.float32(-0.2))
    d = [F * K] + layers
    for i in range(3):
        w["layer_%d" % i] = rng.normal(0, 0.3, (d[i], d[i + 1])).astype(np.float32)
        w["bias_%d" % i] = rng.normal(0, 0.1, (1, d[i + 1])).astype(np.float32)
    X = rng.integers(0, M, (B, F)); Y = rng.choice([1.0, 0.0], B).astype(np.float32)
    loss, z, g = O.wd_loss_grads(X, Y, w)
    tw = {k: torch.tensor(np.asarray(v, np.float64), requires_grad=True) for k, v in w.items()}
    _, c = O.wd_forward(X, w)
    Xt = torch.tensor(X)
    wide = tw["wide_bias"] + tw["wide_linear"][Xt].sum(1)
    for p in range(6):
        wide = wide + tw["wide_cross"][p][torch.tensor(c["buckets"][:, p])]
    h = tw["feature_embeddings"][Xt].reshape(B, -1)
    for i in range(3):
        h = torch.relu(h @ tw["layer_%d" % i] + tw["bias_%d" % i])
    zt = (h @ tw["logits_w"]).reshape(-1) + tw["logits_b"] + wide
    lt = torch.nn.functional.binary_cross_entropy_with_logits(zt, torch.tensor(Y, dtype=torch.float64))
    lt.backward()
    assert abs(float(loss) - float(lt.detach())) < 1e-6
    for k in g:
        a = np.asarray(g[k], np.float64).reshape(-1); b = tw[k].grad.numpy().reshape(-1)
        assert np.abs(a - b).max() <= 2e-6 * (np.abs(b).max() + 1e-30), k

    def splitmix(x):
        x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return x ^ (x >> 31)
    xi, xj = np.array([0, 5, 99999, 4081]), np.array([1, 7, 12, 99998])
    assert O.wd_cross_bucket(xi, xj, 10000).tolist() == [splitmix((int(a) << 32) | int(b)) % 10000 for a, b in zip(xi, xj)]

    gg = rng.normal(0, 1, 100).astype(np.float32)
    w1, a1, z1 = O.ftrl_dense(np.zeros(100, np.float32), np.full(100, 0.1, np.float32), np.zeros(100, np.float32), gg, 0.135)
    wa, aa = O.adagrad_dense(np.zeros(100, np.float32), np.full(100, 0.1, np.float32), gg, 0.135)
    assert np.allclose(w1, wa, rtol=1e-6, atol=1e-9) and np.allclose(a1, aa)
    w2, _, _ = O.ftrl_dense(w1, a1, z1, gg, 0.135, l1=10.0)
    assert (w2 == 0).all()                                   # |z| <= l1 clips to exactly zero


def test_lazy_dense_l2_schedule_is_bit_identical_to_the_dense_schedule():
    """SURVEY.md 7 hard part 2-ii on the CPU: every step the dense schedule moves EVERY row with g + lamda*w (FM.py:124,132);
    the lazy schedule replays a row's missed `g = lamda*w` steps when a batch gathers it (and at the final flush), then applies
    the step with the batch gradient.  Same fp32 operations in the same order per row -> the same bits."""
    rng = np.random.default_rng(17)
    M, K, lr, lam, steps = 40, 8, np.float32(0.1), np.float32(0.05), 12
    w0 = rng.normal(0, 0.3, (M, K)).astype(np.float32)
    a0 = np.full((M, K), 0.1, np.float32)
    touched = [np.unique(rng.integers(0, M, rng.integers(0, 9))) for _ in range(steps)]
    grads = [rng.normal(0, 0.2, (len(t), K)).astype(np.float32) for t in touched]
    # dense schedule
    wd, ad = w0.copy(), a0.copy()
    for t, g in zip(touched, grads):
        G = np.zeros((M, K), np.float32); G[t] = g
        wd, ad = O.adagrad_dense_l2(wd, ad, G, lr, lam)
    # lazy schedule
    wl, al, last = w0.copy(), a0.copy(), np.zeros(M, np.int64)
    for step, (t, g) in enumerate(zip(touched, grads), start=1):
        wl, al, last = O.adagrad_l2_lazy_replay(wl, al, last, t, step - 1, lr, lam)         # bring the gathered rows to step-1
        if len(t):
            wl[t], al[t] = O.adagrad_dense_l2(wl[t], al[t], g, lr, lam)
            last[t] = step
    wl, al, last = O.adagrad_l2_lazy_replay(wl, al, last, None, steps, lr, lam)              # flush
    assert np.array_equal(wd, wl) and np.array_equal(ad, al) and (last == steps).all()
