#!/usr/bin/env python
"""Where the end-to-end step time goes (host pack / H2D / kernels / loss read-back) for bench.py's workload."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench as B
from hhfm_b200.models import OUR
from hhfm_b200.engine import pack_records, Staging

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
n = 1 << 20
hb = B.make_batch(rng, n)
m = OUR(len(B.CTX_CARD), 0, B.FEATURES_M, B.N_USER, B.N_ITEM, B.K_FACTOR, B.LR, B.LAMDA, "AdagradOptimizer", True, False)
for _ in range(3):
    m.partial_fit(hb)
stg = Staging(torch.int32, dev)
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
parts = [hb["X"], hb["F1"], hb["Y"]]
print("cores", os.cpu_count())
print("pack_records ms", t(lambda: pack_records(parts, B.FEATURES_M, stg)))
host, stride = pack_records(parts, B.FEATURES_M, stg)
print("h2d ms", t(lambda: stg.upload(host.numel())), "MB", host.numel() * 4 / 1e6)
idx = stg.upload(host.numel()).view(n, stride)
print("fit_device+loss ms", t(lambda: (m.fit_device(idx, 8, 0, 10), m._read_loss())))
print("partial_fit ms", t(lambda: m.partial_fit(hb)))
up = m._uploader if hasattr(m, "_uploader") else None
for simd in ("0", "1", "0", "1"):
    os.environ["HHFM_PACK_SIMD"] = simd
    print("HHFM_PACK_SIMD=%s partial_fit ms" % simd, t(lambda: m.partial_fit(hb), reps=20))
