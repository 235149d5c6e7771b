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
    # ---- appended later (new random draws only AFTER the ones above, so the earlier vectors stay what they were) ----
    # DeepFM (DFM.py:104-152): tower F*K -> 12 -> 10 -> 6
    layers = [12, 10, 6]
    dw = dict(feature_embeddings=V, feature_bias=rng.uniform(0, 1, (M, 1)).astype(np.float32))
    d = [F * K] + layers
    for i in range(3):
        dw["layer_%d" % i] = rng.normal(0, 0.3, (d[i], d[i + 1])).astype(np.float32)
        dw["bias_%d" % i] = rng.normal(0, 0.1, (1, d[i + 1])).astype(np.float32)
    dw["concat_projection"] = rng.normal(0, 0.3, (F + K + layers[-1], 1)).astype(np.float32)
    dw["concat_bias"] = np.float32(0.01)
    loss, o, g = O.dfm_loss_grads(X, Y * 2 - 1, dw, 0.01)
    for k, v in dw.items():
        if k != "feature_embeddings":
            out["dfm_w_" + k] = np.asarray(v)
    out.update(dfm_loss=loss, dfm_out=o)
    for k, v in g.items():
        out["dfm_g_" + k] = np.asarray(v)
    # Wide&Deep (WDMF.py:51-126 restated): same tower shape, 97 cross buckets
    n_pairs = F * (F - 1) // 2
    ww = dict(feature_embeddings=V, wide_linear=rng.normal(0, 0.1, (M,)).astype(np.float32),
              wide_cross=rng.normal(0, 0.1, (n_pairs, 97)).astype(np.float32), wide_bias=np.float32(0.03),
              logits_w=rng.normal(0, 0.3, (layers[-1], 1)).astype(np.float32), logits_b=np.float32(-0.02))
    for i in range(3):
        ww["layer_%d" % i] = dw["layer_%d" % i]; ww["bias_%d" % i] = dw["bias_%d" % i]
    loss, z, g = O.wd_loss_grads(X, Y, ww)
    out.update(wd_wide_linear=ww["wide_linear"], wd_wide_cross=ww["wide_cross"], wd_logits_w=ww["logits_w"], wd_loss=loss, wd_logit=z,
               wd_buckets=O.wd_wide(X, ww)[1])
    for k, v in g.items():
        out["wd_g_" + k] = np.asarray(v)
    w1, a1, z1 = O.ftrl_dense(ww["wide_linear"], np.full(M, 0.1, np.float32), np.zeros(M, np.float32), g["wide_linear"], 0.135)
    out.update(wd_ftrl_w=w1, wd_ftrl_accum=a1, wd_ftrl_linear=z1)
    # CARS2 (CARS2.py:85-123): D = 20 -> D_c 8, D_p 4, D_q 8
    D = 20
    Dc, Dp, Dq = int(D / 2.5), int(D / 5), int(D / 2.5)
    cw = dict(UI=rng.normal(0, 0.1, (n_user + n_item, D)).astype(np.float32), Context=rng.normal(0, 0.1, (30, Dc)).astype(np.float32),
              W=rng.normal(0, 0.1, (D, Dp, Dc)).astype(np.float32), Z=rng.normal(0, 0.1, (D, Dq, Dc)).astype(np.float32),
              A=rng.normal(0, 0.3, Dp).astype(np.float32), B=rng.normal(0, 0.3, Dq).astype(np.float32))
    Fea = rng.integers(0, 30, B)
    loss, pos, gr = O.cars2_loss_grads(X[:, :2], Fea, Neg[:, :3], cw, 0.001)
    out["cars2_Fea"] = Fea
    for k, v in cw.items():
        out["cars2_w_" + k] = v
    out.update(cars2_loss=np.asarray(loss), cars2_pos=np.asarray(pos))
    for k, v in gr.items():
        out["cars2_g_" + k] = np.asarray(v)
    out["cars2_feedback"] = O.cars2_feedback(X[:, :2], Fea, cw)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **cases())
    print("wrote oracle_vectors.npz")
