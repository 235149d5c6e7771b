"""Parity at BASELINE.json's FULL sizes, where the oracle cannot score every pair in seconds: the bench workloads
(2^20-positive HHFM step on the frappe-10 shape; full-catalog top-N over 10^6 items, K = 128, tp = 100) are checked through
the oracle on what it can reach (every row's scores in chunks, sampled rows of the catalog) plus size-independent properties:
sortedness under the lowest-index tie-break, exactness of every returned score, completeness against an independent fp32 GEMM,
idempotence, and additivity of the loss / gradient over a split of the batch."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import assert_close
from oracle import hhfm_oracle as O

pytestmark = pytest.mark.gpu


def test_full_catalog_topn_at_bench_size(cuda):
    """C = 4096 contexts x N = 10^6 items, K = 128, tp = 100 (the bench's top-N workload at a quarter of its contexts)."""
    from test_gpu_kernels import _topn
    rng = np.random.default_rng(2024)
    n_user, N, K, C, tp = 1024, 1_000_000, 128, 4096, 100
    M = n_user + N
    V = rng.normal(0, 0.01, (M, K)).astype(np.float32)
    V[n_user + 12345] *= 6.0                                              # a large-norm item stresses the filter's error bound
    V[n_user + 500_000:n_user + 500_040] = V[n_user + 77]                 # 40 exact duplicates of one item: ties at every rank they reach
    A = np.stack([rng.integers(0, n_user, C), rng.integers(n_user, M, C)], axis=1)
    info = {}
    ids, sc = _topn(cuda, 0, A, V, None, n_user, N, tp, 0, 0, method="tc", info=info)
    assert info["method"] == "tc"
    # (1) structure: ids in range and unique per row; scores descending, lower index first among equal scores
    assert ids.min() >= 0 and ids.max() < N
    assert (np.sort(ids, axis=1)[:, 1:] != np.sort(ids, axis=1)[:, :-1]).all()
    assert (sc[:, :-1] >= sc[:, 1:]).all()
    tie = sc[:, :-1] == sc[:, 1:]
    assert (ids[:, :-1][tie] < ids[:, 1:][tie]).all()
    assert tie.any(), "the duplicated items should have produced ties"
    # (2) every returned score is the oracle's canonical fp32 dot product, bit for bit (409 600 pairs)
    Q = V[A[:, 0]]
    exact = O.seq_dot(Q[:, None, :], V[n_user + ids])
    assert (exact.view(np.int32) == sc.view(np.int32)).all()
    # (3) sampled rows: the whole list equals the oracle's top_k over the full catalog
    rows = rng.choice(C, 4, replace=False)
    ref = O.dot_topk_scores(Q[rows], V, n_user, N)
    assert (ids[rows] == O.topk_lowest_index(ref, tp)).all()
    # (4) completeness for EVERY row against an independent fp32 GEMM (torch / cuBLAS, no TF32): fewer than tp items may score
    #     above the tp-th returned score; delta covers the summation-order difference between cuBLAS and the canonical order
    assert not torch.backends.cuda.matmul.allow_tf32
    Qd = torch.from_numpy(Q).to(cuda)
    Vd = torch.from_numpy(V[n_user:]).to(cuda)
    kth = torch.from_numpy(sc[:, -1].copy()).to(cuda)
    delta = 4 * K * 1.2e-7 * float(np.abs(Q).max() * np.abs(V).max()) * K ** 0.5
    above = torch.zeros(C, dtype=torch.int64, device=cuda)
    for n0 in range(0, N, 125_000):
        S = Qd @ Vd[n0:n0 + 125_000].T
        above += (S > (kth + delta)[:, None]).sum(dim=1)
    assert int(above.max().item()) < tp, "an item outside some list scores above that list's last entry"
    # (5) idempotence
    ids2, sc2 = _topn(cuda, 0, A, V, None, n_user, N, tp, 0, 0, method="tc")
    assert (ids2 == ids).all() and (sc2.view(np.int32) == sc.view(np.int32)).all()


def test_hhfm_train_pass_at_bench_size(cuda):
    """One fused HHFM forward/backward over 2^20 positives of the frappe-10 shape (bench.py's step, hot-row replicas on):
    the scores of three 65536-row chunks of the full-batch run and the loss / gradient of one chunk against the oracle, and
    additivity of loss and gradient over a split of the whole batch."""
    from test_gpu_kernels import _pairrank_train
    import bench as Bm
    rng = np.random.default_rng(99)
    B = 1 << 20
    hb = Bm.make_batch(rng, B)
    M, K = Bm.FEATURES_M, Bm.K_FACTOR
    V = rng.normal(0, 0.05, (M, K)).astype(np.float32)
    Pos, Fea, Neg = hb["X"], hb["F1"], hb["Y"]
    counts = np.bincount(np.concatenate([Pos.reshape(-1), Fea.reshape(-1), Neg.reshape(-1)]), minlength=M)
    hot_rows = np.argsort(-counts)[:1024]
    got = _pairrank_train(cuda, V, Pos, Fea, None, Neg, (0, 0, 0), hot_rows=hot_rows)
    # oracle on three 65536-row chunks (a few seconds each): every score of the full-batch run inside them, and loss / gradient
    # of the middle chunk run as its own batch; additivity (below) carries that to the whole batch
    step = 1 << 16
    for r0 in (0, 7 * step, B - step):
        sl = slice(r0, r0 + step)
        l, pos, neg, g = O.pairrank_loss_grads(V, Pos[sl], Neg[sl], Fea[sl], None, (0, 0, 0), 0.0)
        assert_close(got["pos"][sl], pos, what="pos rows %d.." % r0)
        assert_close(got["neg"][sl], neg, what="neg rows %d.." % r0)
        if r0 == 7 * step:
            c = _pairrank_train(cuda, V, Pos[sl], Fea[sl], None, Neg[sl], (0, 0, 0), hot_rows=hot_rows)
            assert_close(c["loss"], l, what="loss of chunk 7")
            assert_close(c["gV"], g, rtol=2e-5, what="gV of chunk 7")
    # rows hit by up to ~5e5 samples: the halves and the whole add the same fp32 terms in different orders (atomics, replicas),
    # so the gradient identity is held to 1e-4 of max(|ref|, rms) instead of 1e-5
    # additivity: loss and gradient of the batch = those of its two halves
    h = B // 2
    a = _pairrank_train(cuda, V, Pos[:h], Fea[:h], None, Neg[:h], (0, 0, 0), hot_rows=hot_rows)
    b = _pairrank_train(cuda, V, Pos[h:], Fea[h:], None, Neg[h:], (0, 0, 0), hot_rows=hot_rows)
    assert_close(got["loss"], a["loss"] + b["loss"], what="loss additivity")
    assert_close(got["gV"], a["gV"] + b["gV"], rtol=1e-4, what="gradient additivity")
    assert set(got["touched"].tolist()) == set(np.unique(np.concatenate([a["touched"], b["touched"]])).tolist())
