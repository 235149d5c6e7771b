#!/usr/bin/env python
"""Stage times of the tensor-core top-N on ONE item shard (what a rank of the item-sharded evaluator runs):
    python scripts/topn_stage_times.py [--contexts 16384] [--items 125000] [--tp 100] [--local-tp 0]
Run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel list; by itself it prints the time of the whole call."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--contexts", type=int, default=16384)
    ap.add_argument("--items", type=int, default=125000)
    ap.add_argument("--tp", type=int, default=100)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    from hhfm_b200.engine import TopN
    dev = torch.device("cuda", 0)
    C, N, K, tp, n_user = args.contexts, args.items, 128, args.tp, 4096
    g = torch.Generator(device="cpu").manual_seed(777)
    V = torch.empty(n_user + N, K).normal_(0, 0.01, generator=g).to(dev)
    A = torch.stack([torch.randint(0, n_user, (C,), generator=g), torch.full((C,), n_user, dtype=torch.int64)], 1).to(torch.int32)
    t = TopN(dev, max_workspace_bytes=6 << 30)
    A_dev, stride = t.upload_rows(A.numpy(), n_user + N)

    def once():
        return t.topk(0, A_dev, stride, 0, 0, (0, 0, 0), V, None, n_user, N, tp, return_scores=True, method="tc", version=1)

    for _ in range(2):
        once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        once()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    print(json.dumps({"contexts": C, "items": N, "tp": tp, "ms": ms, "gemm_ideal_ms_at_1.6PF": 2.0 * K * C * N / 1.6e15 * 1e3,
                      "overflow_rows": t.last_overflow_rows}))


if __name__ == "__main__":
    main()
