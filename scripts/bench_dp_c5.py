#!/usr/bin/env python
"""Data-parallel training at the scaled c5 shape (tables too large for a dense all-reduce): HHFM, M = 10^7 ids,
K = 128, B = 2^20 positives PER RANK, lamda = 0, coalesced-sparse row exchange (`enable_data_parallel(sparse=True)`: touched
rows packed, all-gathered, added in rank order, touched-row Adagrad on the union).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_dp_c5.py [--steps 6]

Prints one JSON line on rank 0: ms per step (max over ranks), samples/s over the job, bytes exchanged per rank and step, and
the check that the replicas are bit-identical after the timed steps.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from bench_models import zipf_ids
    from hhfm_b200.engine import Staging, pack_records
    from hhfm_b200.models import OUR
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(55 + rank)
    B, K, NG = 1 << 20, 128, 10
    n_user, n_item, n_ctx_ids = 4_000_000, 1_000_000, 5_000_000
    M = n_user + n_item + n_ctx_ids
    per_ctx = n_ctx_ids // 8
    recs = []
    for b in range(2):
        X = np.stack([zipf_ids(rng, n_user, B), n_user + zipf_ids(rng, n_item, B)], 1).astype(np.int64)
        base = n_user + n_item
        cols = []
        for c in range(8):
            cols.append(base + rng.integers(0, per_ctx, B)); base += per_ctx
        F1 = np.stack(cols, 1).astype(np.int64)
        Y = (n_user + rng.integers(0, n_item, (B, NG))).astype(np.int64)
        stg = Staging(torch.int32, dev)
        host, stride = pack_records([X, F1, Y], M, stg)
        recs.append(stg.upload(host.numel()).view(B, stride).clone())
        del stg
    m = OUR(8, 0, M, n_user, n_item, K, 0.1, 0.0, "AdagradOptimizer", True, False)
    if world > 1:
        m.enable_data_parallel(sparse=True)
    for i in range(2):
        m.fit_device(recs[i % 2], 8, 0, NG)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        m.fit_device(recs[i % 2], 8, 0, NG)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    uniq = int(m._touch.count.item())
    same = True
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        V = m.weights["feature_embeddings"]
        chk = torch.stack([V[::4097].double().sum(), V.view(torch.int32)[::1021].long().sum().double()])
        g = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(g, chk)
        same = all(bool(torch.equal(g[0], x)) for x in g)
    if rank == 0:
        print(json.dumps({"workload": "HHFM c5 shape, data parallel, coalesced-sparse row exchange", "n_gpus": world,
                          "ms_per_step": ms, "samples_per_s": world * B / ms * 1e3, "rows_in_union_per_step": uniq,
                          "exchange_bytes_per_rank_per_step": int(uniq * (4 * K + 4)), "replica_checksums_identical": same}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
