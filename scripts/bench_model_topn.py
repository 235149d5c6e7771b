#!/usr/bin/env python
"""AFM / DeepFM full-catalog evaluation (AFM.topk AFM.py:209-246, DeepFM.topk DFM.py:219-231) on one B200: the item-separable
scorers (csrc/afm_topn.cu, hhfm_dfm_topn_scores) against sending every (row, item) pair through the forward kernels.
frappe-10 shape: F = 10, K = 64, N = 4082 items.

    python scripts/bench_model_topn.py [--model afm|dfm] [--contexts 2048] [--reps 5]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--contexts", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--model", choices=["afm", "dfm"], default="afm")
    args = ap.parse_args()
    import torch
    from bench_models import frappe_rows
    from hhfm_b200.models import AFM
    rng = np.random.default_rng(3)
    X, M, n_user, n_item = frappe_rows(rng, args.contexts)
    K = 64
    if args.model == "afm":
        m = AFM(n_user, n_item, M, 1, [K, K], "relu", 0.1, 100.0, [1, 1], "AdagradOptimizer", 0.999, 10)
    else:
        from hhfm_b200.models import DeepFM
        m = DeepFM(n_user, n_item, M, 10, K, [150, 200, 150], "relu", 0.01, 0, 0.01)
    out = {}
    lists = {}
    for name, env in (("separable", "1"), ("full_forward", "0")):
        os.environ["HHFM_AFM_TOPN_SEPARABLE" if args.model == "afm" else "HHFM_DFM_TOPN_SEPARABLE"] = env
        lists[name] = m.topk(X, 20)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            m.topk(X, 20)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / args.reps * 1e3
        out[name] = {"ms_per_call": ms, "pairs_per_s": args.contexts * n_item / ms * 1e3}
    out["lists_equal_frac"] = float((lists["separable"] == lists["full_forward"]).mean())
    out["config"] = ("AFM" if args.model == "afm" else "DeepFM (640-150-200-150)") + ".topk, frappe-10 shape (F=10, K=64, N=%d items), C=%d context rows, tp=20, host rows in / lists out" % (n_item, args.contexts)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
