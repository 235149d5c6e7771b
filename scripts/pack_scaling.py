#!/usr/bin/env python
"""Where the end-to-end step goes (bench.py `e2e`): host packing + H2D of the int64 feed (hhfm_pack_upload_records) against the
number of packer threads (one child process per count: the pool is created once per process), the raw H2D copy of the wire
records, and the whole partial_fit / partial_fit_async call.
    python scripts/pack_scaling.py"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    import bench
    from hhfm_b200 import engine
    from hhfm_b200.models import OUR
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    B = 1 << 20
    batches = [bench.make_batch(rng, B) for _ in range(3)]
    out = {"threads": engine._NTHREADS}
    up = engine.RecordUploader(dev)
    for i in range(3):
        up.upload([batches[i % 3][k] for k in ("X", "F1", "Y")], bench.FEATURES_M)
    torch.cuda.synchronize()
    n = 10
    t0 = time.perf_counter()
    for i in range(n):
        up.upload([batches[i % 3][k] for k in ("X", "F1", "Y")], bench.FEATURES_M)
    host_ms = (time.perf_counter() - t0) / n * 1e3          # host time per call (the last chunk's copy may still be in flight)
    torch.cuda.synchronize()
    out["upload_host_ms"] = host_ms
    m = OUR(8, 0, bench.FEATURES_M, bench.N_USER, bench.N_ITEM, bench.K_FACTOR, bench.LR, bench.LAMDA, "AdagradOptimizer", True, False)
    for i in range(3):
        m.partial_fit(batches[i % 3])
    t0 = time.perf_counter()
    for i in range(n):
        m.partial_fit(batches[i % 3])
    out["partial_fit_ms"] = (time.perf_counter() - t0) / n * 1e3
    t0 = time.perf_counter()
    prev = None
    for i in range(n):
        h = m.partial_fit_async(batches[i % 3])
        if prev is not None:
            prev.result()
        prev = h
    prev.result()
    out["partial_fit_async_ms"] = (time.perf_counter() - t0) / n * 1e3
    idx = up.upload([batches[0][k] for k in ("X", "F1", "Y")], bench.FEATURES_M)[0].clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        m.fit_device(idx, 8, 0, bench.NG)
    host_enqueue = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    out["fit_device_host_enqueue_ms"] = host_enqueue
    print(json.dumps(out))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        return child()
    import torch
    res = {"cores": os.cpu_count(), "runs": []}
    for nt in (1, 2, 4, 8, 16, 32):
        if nt > 2 * res["cores"]:
            break
        env = dict(os.environ, HHFM_PACK_THREADS=str(nt))
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
        line = [l for l in p.stdout.splitlines() if l.startswith("{")]
        res["runs"].append(json.loads(line[-1]) if line else {"threads": nt, "error": p.stderr[-300:]})
    dev = torch.device("cuda", 0)
    B = 1 << 20
    h = torch.empty(B * 20 * 2, dtype=torch.uint8, pin_memory=True)
    d = torch.empty_like(h, device=dev)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    res["h2d_42MB_ms"] = (time.perf_counter() - t0) / 10 * 1e3
    print(json.dumps(res))


if __name__ == "__main__":
    main()
