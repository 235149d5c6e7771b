#!/usr/bin/env python
"""Wide&Deep (WDMF.py:51-126 restated) full-batch step on one B200: frappe-sized batch (2 x 88 571 rows, F = 10), embeddings 128,
DNN 1024-512-256, FTRL wide half + Adagrad deep half.  One `fit_device` = one of the 500 steps of the reference's `partial_fit`.

    python scripts/bench_wd.py [--rows 177142] [--steps 5]
"""
from __future__ import annotations

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
    ap.add_argument("--rows", type=int, default=177142)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import torch
    from bench_models import frappe_rows, timed
    from hhfm_b200.models import WD
    rng = np.random.default_rng(8)
    X, M, n_user, n_item = frappe_rows(rng, args.rows)
    y = torch.from_numpy(rng.choice([1.0, 0.0], args.rows).astype(np.float32)).cuda()
    m = WD(10, n_user, n_item)
    idx = torch.from_numpy(X).cuda()
    ms = timed(lambda i: m.fit_device(idx, y), args.steps, 2)
    dims = [10 * 128, 1024, 512, 256, 1]
    flops = 3 * 2 * sum(dims[i] * dims[i + 1] for i in range(4))
    print(json.dumps({"config": "WD frappe-10 full batch (%d rows), embeddings 128, DNN 1024-512-256" % args.rows, "ms_per_step": ms,
                      "samples_per_s": args.rows / ms * 1e3, "algorithmic_flops_per_sample": flops,
                      "tflops": args.rows * flops / ms / 1e9, "loss": m._read_loss()}))


if __name__ == "__main__":
    main()
