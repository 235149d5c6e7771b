"""GPU, multi-process: data-parallel training and item-sharded evaluation over NCCL.  Runs with however many GPUs the
box has (skips the multi-GPU part on a 1-GPU box but still checks the world-size-1 path)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from hhfm_b200 import dist as hd
        from hhfm_b200.models import OUR
        from oracle import hhfm_oracle as O
        rng = np.random.default_rng(0)
        n_user, n_item, M, K, fc, B = 100, 301, 500, 64, 4, 4096
        X = np.stack([rng.integers(0, n_user, B), n_user + rng.integers(0, n_item, B)], axis=1)
        F1 = rng.integers(n_user + n_item, M, (B, fc)); Y = n_user + rng.integers(0, n_item, (B, 10))
        lo, hi = hd.shard_range(B, rank, world)
        finals = {}
        results = []
        for mode in ("auto", False):           # fused fold + multimem all-reduce + optimizer kernel / NCCL all-reduce
            model = OUR(fc, 0, M, n_user, n_item, K, 0.1, 0.01, 'AdagradOptimizer', True, False)
            model.enable_data_parallel(p2p=mode); model.enable_item_sharding()
            if world > 1 and mode == "auto":
                assert model._dpx is not None, "symmetric memory should be available on one NVLink box"
            V = model.get_weights()["feature_embeddings"].copy()
            loss = model.partial_fit({"X": X[lo:hi], "F1": F1[lo:hi], "Y": Y[lo:hi]})
            loss_ref, _, _, dV = O.pairrank_loss_grads(V, X, Y, F1, None, (0, 0, 0), 0.01)
            V1, _ = O.adagrad_dense(V, np.full_like(V, 0.1), dV, 0.1)
            got = model.get_weights()["feature_embeddings"]
            ok_loss = abs(loss - loss_ref) <= 2e-5 * abs(loss_ref)
            ok_w = bool(np.all(np.abs(got - V1) <= 2e-5 * np.maximum(np.abs(V1 - V), np.sqrt(np.mean((V1 - V) ** 2)))))
            A = np.concatenate([X[:200], F1[:200]], axis=1)
            ids = model.topk(A, 20)
            want = O.topk_lowest_index(O.hhfm_topk_scores(A, got, n_user, n_item, fc, 0), 20)
            # the other decomposition of the evaluator: context rows sharded, whole catalog per rank, lists concatenated
            eg, model._eval_group = model._eval_group, None
            model.enable_context_sharding()
            ids_ctx = model.topk(A[:197], 20)                    # 197 rows: uneven shards
            model._eval_ctx_group, model._eval_group = None, eg
            ok_ctx = ids_ctx.shape == (197, 20) and bool((ids_ctx == want[:197]).all())
            # the HR / NDCG walk sharded by context rows (device walk on this rank's rows + all-reduce of the sums over NCCL)
            from hhfm_b200 import engine
            tgt = rng.integers(0, n_item, 197).astype(np.int32); inpf = (rng.random(197) < 0.1).astype(np.uint8)
            full = engine.metrics_walk(torch.from_numpy(want[:197].astype(np.int32)).cuda(), torch.from_numpy(tgt).cuda(),
                                       torch.from_numpy(inpf).cuda(), 5).cpu().numpy()
            r0, r1 = hd.shard_range(197, rank, world)
            mine = engine.metrics_walk(torch.from_numpy(ids_ctx[r0:r1].astype(np.int32)).cuda(), torch.from_numpy(tgt[r0:r1]).cuda(),
                                       torch.from_numpy(inpf[r0:r1]).cuda(), 5).cpu().numpy()
            ok_ctx = ok_ctx and bool(np.allclose(hd.allreduce_metrics(mine), engine.metrics_from_codes(full), rtol=1e-12, atol=0))
            # three more steps (both arena buffers get reused), then the replicas must still be identical on every rank
            for step in range(3):
                r2 = np.random.default_rng(100 + step)
                Xs = np.stack([r2.integers(0, n_user, B), n_user + r2.integers(0, n_item, B)], axis=1)
                Fs = r2.integers(n_user + n_item, M, (B, fc)); Ys = n_user + r2.integers(0, n_item, (B, 10))
                model.partial_fit({"X": Xs[lo:hi], "F1": Fs[lo:hi], "Y": Ys[lo:hi]})
            w = model.weights["feature_embeddings"]
            gathered = [torch.empty_like(w) for _ in range(world)]
            dist.all_gather(gathered, w)
            same = all(bool(torch.equal(gathered[0], g)) for g in gathered)
            finals[mode] = w.cpu().numpy().copy()
            results.append((ok_loss, ok_w, bool((ids == want).all()) and ok_ctx, same))
        d = np.abs(finals["auto"] - finals[False])
        close = bool(np.all(d <= 1e-4 * np.maximum(np.abs(finals[False]), np.sqrt(np.mean(finals[False] ** 2)))))
        ret[rank] = tuple(results) + (close,)
    finally:
        dist.destroy_process_group()


def test_data_parallel_step_and_sharded_topk(cuda):
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 2)
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        assert ret[r] == ((True, True, True, True), (True, True, True, True), True), (r, ret[r])


def _sparse_worker(rank, world, port, ret):
    """Coalesced-sparse gradient exchange: lamda = 0 models, rows sharded, every rank ends on the weights of the
    single-process step on the full batch; replicas bit-identical; rows nobody touched do not move."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from hhfm_b200 import dist as hd
        from hhfm_b200.models import FM, OUR
        from oracle import hhfm_oracle as O
        rng = np.random.default_rng(5)
        n_user, n_item, M, K, B = 300, 700, 1200, 64, 4096
        out = []
        # ---- HHFM (pair ranking, no bias), lamda = 0 ----
        fc = 4
        X = np.stack([rng.integers(0, n_user, B), n_user + rng.integers(0, n_item, B)], axis=1)
        F1 = rng.integers(n_user + n_item, M - 100, (B, fc))
        Y = n_user + rng.integers(0, n_item, (B, 10))
        lo, hi = hd.shard_range(B, rank, world)
        model = OUR(fc, 0, M, n_user, n_item, K, 0.1, 0.0, 'AdagradOptimizer', True, False)
        model.enable_data_parallel(sparse=True)
        assert model._dp_sparse == (world > 1)
        V = model.get_weights()["feature_embeddings"].copy()
        loss = model.partial_fit({"X": X[lo:hi], "F1": F1[lo:hi], "Y": Y[lo:hi]})
        loss_ref, _, _, dV = O.pairrank_loss_grads(V, X, Y, F1, None, (0, 0, 0), 0.0)
        V1, _ = O.adagrad_dense(V, np.full_like(V, 0.1), dV, 0.1)
        got = model.get_weights()["feature_embeddings"]
        upd = V1 - V
        ok = abs(loss - loss_ref) <= 2e-5 * abs(loss_ref)
        ok = ok and bool(np.all(np.abs(got - V1) <= 1e-4 * np.maximum(np.abs(upd), np.sqrt(np.mean(upd ** 2)))))
        ok = ok and bool((got[M - 100:] == V[M - 100:]).all())                          # untouched rows
        for step in range(3):
            r2 = np.random.default_rng(200 + step)
            Xs = np.stack([r2.integers(0, n_user, B), n_user + r2.integers(0, n_item, B)], axis=1)
            Ys = n_user + r2.integers(0, n_item, (B, 10)); Fs = r2.integers(n_user + n_item, M - 100, (B, fc))
            model.partial_fit({"X": Xs[lo:hi], "F1": Fs[lo:hi], "Y": Ys[lo:hi]})
        w = model.weights["feature_embeddings"]
        gathered = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(gathered, w)
        out.append((bool(ok), all(bool(torch.equal(gathered[0], g)) for g in gathered)))
        # ---- FM (pointwise, feature bias + scalar bias) ----
        F = 6
        Xf = np.concatenate([X, rng.integers(n_user + n_item, M - 100, (B, F - 2))], axis=1)
        Yf = rng.choice([1.0, 0.0], (B, 1)).astype(np.float32)
        fm = FM(F, M, n_user, n_item, K, 0.1, 0.0, 1, 'AdagradOptimizer', 0, 0)
        fm.enable_data_parallel(sparse=True)
        w0 = fm.get_weights()
        loss = fm.partial_fit({"X": Xf[lo:hi], "Y": Yf[lo:hi]})
        loss_ref, _, dV, db, db0, _ = O.fm_loss_grads(Xf, Yf, w0["feature_embeddings"], w0["feature_bias"], w0["bias"], 0.0)
        V1, _ = O.adagrad_dense(w0["feature_embeddings"], np.full_like(w0["feature_embeddings"], 0.1), dV, 0.1)
        b1, _ = O.adagrad_dense(w0["feature_bias"], np.full_like(w0["feature_bias"], 0.1), np.asarray(db).reshape(-1, 1), 0.1)
        got = fm.get_weights()
        upd = V1 - w0["feature_embeddings"]; updb = b1 - w0["feature_bias"]
        ok = abs(loss - loss_ref) <= 2e-5 * abs(loss_ref)
        ok = ok and bool(np.all(np.abs(got["feature_embeddings"] - V1) <= 1e-4 * np.maximum(np.abs(upd), np.sqrt(np.mean(upd ** 2)))))
        ok = ok and bool(np.all(np.abs(got["feature_bias"] - b1) <= 1e-4 * np.maximum(np.abs(updb), np.sqrt(np.mean(updb ** 2)))))
        ok = ok and bool((got["feature_embeddings"][M - 100:] == w0["feature_embeddings"][M - 100:]).all())
        for step in range(2):
            fm.partial_fit({"X": Xf[lo:hi][::-1].copy(), "Y": Yf[lo:hi][::-1].copy()})
        same = True
        for k in ("feature_embeddings", "feature_bias"):
            w = fm.weights[k]
            gathered = [torch.empty_like(w) for _ in range(world)]
            dist.all_gather(gathered, w)
            same = same and all(bool(torch.equal(gathered[0], g)) for g in gathered)
        out.append((bool(ok), same))
        ret[rank] = tuple(out)
    finally:
        dist.destroy_process_group()


def test_data_parallel_coalesced_sparse_exchange(cuda):
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 2)
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_sparse_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        assert ret[r] == ((True, True), (True, True)), (r, ret[r])
