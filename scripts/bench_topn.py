"""Explore full-catalog top-N throughput: exact SIMT path vs tcgen05 filter path, with per-stage CUDA-event timings."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hhfm_b200 import _lib
from hhfm_b200.engine import TopN, cur_stream, ptr

ap = argparse.ArgumentParser()
ap.add_argument("--C", type=int, default=4096); ap.add_argument("--N", type=int, default=1000000)
ap.add_argument("--K", type=int, default=128); ap.add_argument("--tp", type=int, default=100)
ap.add_argument("--kind", type=int, default=0); ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--exact", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(1)
n_user = 1024
M = n_user + a.N + 64
V = torch.empty(M, a.K).normal_(0, 0.01, generator=g).to(dev)
bias = torch.empty(M, 1).normal_(0, 0.01, generator=g).to(dev)
F = 2 if a.kind == 0 else 10
A = torch.stack([torch.randint(0, n_user, (a.C,)), torch.randint(n_user, n_user + a.N, (a.C,))] +
                [torch.randint(n_user + a.N, M, (a.C,)) for _ in range(F - 2)], dim=1).to(torch.int32)
t = TopN(dev, max_workspace_bytes=4 << 30)
A_dev, stride = t.upload_rows(A.numpy(), M)
def run(method):
    return t.topk(a.kind, A_dev, stride, F - 2 if a.kind else 0, 0, (0, 0, 0), V, bias if a.kind == 1 else None, n_user, a.N, a.tp,
                  return_scores=True, method=method, version=1)
res = {}
for method in (["tc", "exact"] if a.exact else ["tc"]):
    ids, sc = run(method); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        ids, sc = run(method)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    res[method] = {"ms": ms, "pairs_per_s": a.C * a.N / ms * 1e3, "overflow_rows": t.last_overflow_rows}
    res[method + "_ids"] = ids
if a.exact:
    res["identical"] = bool((res["tc_ids"] == res["exact_ids"]).all())
print(json.dumps({k: v for k, v in res.items() if not k.endswith("_ids")} | {"C": a.C, "N": a.N, "K": a.K, "tp": a.tp, "kind": a.kind}))
