"""Golden vectors of the CPU oracle itself (SURVEY.md 8c "golden vectors to create"): tiny seeded cases with injected
weights -> out / loss / gradients / post-step weights / top-k lists.  They freeze the oracle (a later edit that changes
its arithmetic fails tests/test_oracle.py) and give the GPU tests fixed input/output pairs.

    python tests/golden/make_oracle_goldens.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import hhfm_oracle as O  # noqa: E402


def cases():
    rng = np.random.Generator(np.random.PCG64(2016))
    out = {}
    n_user, n_item, M, K, B, F, NG = 20, 40, 80, 16, 64, 6, 10
    V = rng.normal(0, 0.01, (M, K)).astype(np.float32)
    b = rng.normal(0, 0.01, (M, 1)).astype(np.float32)
    X = np.concatenate([rng.integers(0, n_user, (B, 1)), rng.integers(n_user, n_user + n_item, (B, 1)),
                        rng.integers(n_user + n_item, M, (B, F - 2))], axis=1)
    X[:, 5] = X[:, 4]                                       # duplicate id inside every row
    Y = rng.choice([1.0, 0.0], (B, 1)).astype(np.float32)
    out.update(V=V, b=b, X=X, Y=Y, dims=np.array([n_user, n_item, M, K, B, F, NG]))
    loss, o, dV, db, db0, rows = O.fm_loss_grads(X, Y, V, b, np.float32(0.05), 0.1)
    V1, acc1 = O.adagrad_dense(V, np.full_like(V, 0.1), dV, 0.1)
    out.update(fm_loss=loss, fm_out=o, fm_dV=dV, fm_db=db, fm_db0=db0, fm_V1=V1, fm_acc1=acc1)
    out["fm_top20"] = O.topk_lowest_index(O.fm_topk_scores(X[:16], V, b, n_user, n_item), 20)
    Neg = rng.integers(n_user, n_user + n_item, (B, NG)); Neg[::5] = Neg[::5, :1]   # all-same negatives -> ties
    out["Neg"] = Neg
    for tag, pools in (("sum", (0, 0, 0)), ("max", (1, 1, 1)), ("mean", (2, 2, 2))):
        loss, pos, neg, dV = O.pairrank_loss_grads(V, X[:, :2], Neg, X[:, 2:4], X[:, 4:6], pools, 0.01)
        out.update({"hhfm_%s_loss" % tag: loss, "hhfm_%s_pos" % tag: pos, "hhfm_%s_neg" % tag: neg, "hhfm_%s_dV" % tag: dV})
        out["hhfm_%s_top20" % tag] = O.topk_lowest_index(O.hhfm_topk_scores(X[:16], V, n_user, n_item, 2, 2, pools), 20)
    loss, pos, neg, dV = O.pairrank_loss_grads(V, X[:, :2], Neg, None, None, (0, 0, 0), 0.1)
    out.update(bpr_loss=loss, bpr_dV=dV)
    loss, o, dV = O.mf_loss_grads(X[:, :2], Y * 2 - 1, V, 0.01)
    out.update(mf_loss=loss, mf_out=o, mf_dV=dV)
    w = dict(feature_embeddings=V, feature_bias=b, bias=np.float32(0.02),
             attention_W=rng.normal(0, 0.25, (K, K)).astype(np.float32), attention_b=rng.normal(0, 0.25, (1, K)).astype(np.float32),
             attention_p=rng.normal(0, 1, (K,)).astype(np.float32), prediction=np.ones((K, 1), np.float32))
    loss, o, g = O.afm_loss_grads(X, Y * 2 - 1, w, 100.0)
    out.update(afm_W=w["attention_W"], afm_b=w["attention_b"], afm_p=w["attention_p"], afm_loss=loss, afm_out=o)
    for k, v in g.items():
        out["afm_g_" + k] = np.asarray(v)
    out["afm_top20"] = O.topk_lowest_index(O.afm_topk_scores(X[:8], w, n_user, n_item), 20)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **cases())
    print("wrote oracle_vectors.npz")
