#!/usr/bin/env python
"""Per-model training-step throughput on one B200 with device-resident batches (the BASELINE.json configs that are
not bench.py's headline line): FM c1 (frappe), FM c5 (scaled, HBM-bound), BPR c4, AFM c3, DeepFM.

    python scripts/bench_models.py [--only fm_c1,fm_c5,bpr_c4,afm_c3,dfm] [--steps 20] [--json out.json]

Each line: samples/s from CUDA events over `steps` whole steps (kernel + scatter fold + optimizer + loss reduction), and
the roofline fraction from SURVEY.md 8(d)'s algorithmic bytes (HBM-bound models) or flops (AFM / DeepFM, fp32 SIMT).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FP32_SIMT_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # 148 SMs x 128 FMA lanes x 2 flop x max SM clock = 74.4 TF


def tensor_peak_tf():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p)).get("bf16_tflops", 1626.2))
    return 1590.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"])
    return 6650.0


def _traffic(kernel):
    """DRAM bytes per launch of `kernel` and the ncu capture they come from (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p)).get(kernel) or {}
    return d.get("dram_bytes_per_launch"), d.get("capture")


_L2_PEAK = {}


def l2_peak(dev):
    """L2 -> SM read bandwidth measured in this process (hhfm_l2_read_sweep over an L2-resident 48 MB buffer), GB/s: the
    roofline denominator of the shapes whose table lives in L2 (an HBM fraction above 1 says nothing there)."""
    if "v" not in _L2_PEAK:
        import torch
        from hhfm_b200 import _lib
        from hhfm_b200.engine import cur_stream, ptr
        n = (48 << 20) // 4
        buf = torch.ones(n, dtype=torch.float32, device=dev)
        sink = torch.zeros(1, dtype=torch.float32, device=dev)
        _lib.call("hhfm_l2_read_sweep", ptr(buf), n, 4, ptr(sink), cur_stream())
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.call("hhfm_l2_read_sweep", ptr(buf), n, 20, ptr(sink), cur_stream())
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 4.0 * n * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        _L2_PEAK["v"] = best
    return _L2_PEAK["v"]


def l2_roofline(B, algo, ms, dev):
    ach = B * algo / ms / 1e6
    pk = l2_peak(dev)
    return {"bound": "l2", "achieved": ach, "peak": pk, "unit": "GB/s", "frac": ach / pk,
            "peak_source": "hhfm_l2_read_sweep, measured in this run", "frac_of_hbm_peak_by_algorithmic_bytes": ach / peaks(),
            "note": "the table and the hot-row replicas are L2 / L1 resident at this shape: only the id records stream from HBM, so "
                    "the algorithmic bytes are compared with the measured L2 read bandwidth (a part of them is served by L1)"}


def zipf_ids(rng, n, size, a=1.1):
    # inverse-CDF sampling of a truncated Zipf without materialising n probabilities for huge n
    if n <= 1 << 16:
        p = 1.0 / np.arange(1, n + 1) ** a
        p /= p.sum()
        return rng.choice(n, size=size, p=p)
    u = rng.random(size)
    # continuous approximation: P(X <= x) ~ (1 - x^(1-a)) / (1 - n^(1-a))
    x = (1.0 - u * (1.0 - float(n) ** (1.0 - a))) ** (1.0 / (1.0 - a))
    return np.minimum(n - 1, np.floor(x - 1.0 + 1e-9).astype(np.int64).clip(0))


def frappe_rows(rng, B):
    n_user, n_item, ctx = 957, 4082, (7, 2, 3, 2, 9, 80, 233, 7)
    cols = [zipf_ids(rng, n_user, B), n_user + zipf_ids(rng, n_item, B)]
    base = n_user + n_item
    for c in ctx:
        cols.append(base + rng.integers(0, c, B))
        base += c
    return np.stack(cols, 1).astype(np.int32), base, n_user, n_item


def timed(fn, steps, warmup):
    import torch
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run_fm_c1(args, dev):
    import torch
    from hhfm_b200.models import FM
    rng = np.random.default_rng(1)
    B = 1 << 20
    batches = []
    for _ in range(4):
        X, M, n_user, n_item = frappe_rows(rng, B)
        batches.append((torch.from_numpy(X).to(dev), torch.from_numpy(rng.choice([1.0, 0.0], B).astype(np.float32)).to(dev)))
    m = FM(10, M, n_user, n_item, 64, 0.1, 0.1, 1, "AdagradOptimizer", 0, 0)
    ms = timed(lambda i: m.fit_device(*batches[i % 4]), args.steps, 5)
    algo = 2648 + 2600
    return {"config": "c1 FM frappe-10 (F=10, M=%d, K=64, dense-L2 Adagrad), B=2^20" % M, "ms_per_step": ms,
            "samples_per_s": B / ms * 1e3, "algorithmic_bytes_per_sample": algo,
            "roofline": l2_roofline(B, algo, ms, dev)}


def run_fm_c5(args, dev):
    import torch
    from hhfm_b200.models import FM
    rng = np.random.default_rng(5)
    B, K = 1 << 20, 128
    n_user, n_item, n_ctx_ids = 4_000_000, 1_000_000, 5_000_000
    M = n_user + n_item + n_ctx_ids
    per_ctx = n_ctx_ids // 8
    batches = []
    for _ in range(3):
        cols = [zipf_ids(rng, n_user, B), n_user + zipf_ids(rng, n_item, B)]
        base = n_user + n_item
        for c in range(8):
            cols.append(base + rng.integers(0, per_ctx, B))
            base += per_ctx
        X = np.stack(cols, 1).astype(np.int32)
        batches.append((torch.from_numpy(X).to(dev), torch.from_numpy(rng.choice([1.0, 0.0], B).astype(np.float32)).to(dev)))
    uniq = int(np.unique(batches[0][0].cpu().numpy()).size)
    m = FM(10, M, n_user, n_item, K, 0.1, 0.0, 1, "AdagradOptimizer", 0, 0)      # lamda = 0: sparse (IndexedSlices) update
    ms = timed(lambda i: m.fit_device(*batches[i % 3]), args.steps, 3)
    F = 10
    algo = (4 * F + 4 * F * K + 4 * F + 4 + 4) + (4 * F * K + 4 * F) + 20 * (K + 1) * uniq / B
    return {"config": "c5 scaled FM (F=10, M=10^7, K=128, sparse Adagrad rows), B=2^20, %d unique rows/step" % uniq,
            "ms_per_step": ms, "samples_per_s": B / ms * 1e3, "algorithmic_bytes_per_sample": algo,
            "roofline": {"bound": "hbm", "achieved_gbs": B * algo / ms / 1e6, "peak_gbs": peaks(),
                         "frac": B * algo / ms / 1e6 / peaks(), "frac_of_nominal_8000_gbs": B * algo / ms / 1e6 / 8000.0}}


def run_hhfm_c5(args, dev):
    """OurModel7 (HHFM) at the scaled c5 shape: M = 10^7 ids (4 M users, 1 M items, 8 context columns of
    625 000 values), K = 128, NG = 10, B = 2^20 positives, sparse (lamda = 0) Adagrad rows.  The 5.1 GB table lives in HBM,
    so this is the configuration where the headline model is HBM-bound.  Algorithmic bytes per positive: ids 80 + gather
    (F+NG) = 20 rows x 512 B + scatter 11 rows x 512 B (user, item+, 8 context rows, the arg-max negative) + optimizer
    20*K B per unique touched row."""
    import torch
    from hhfm_b200.models import OUR
    from hhfm_b200.engine import Staging, pack_records
    rng = np.random.default_rng(55)
    B, K, NG = 1 << 20, 128, 10
    n_user, n_item, n_ctx_ids = 4_000_000, 1_000_000, 5_000_000
    M = n_user + n_item + n_ctx_ids
    per_ctx = n_ctx_ids // 8
    recs, uniq = [], 0
    for b in range(3):
        X = np.stack([zipf_ids(rng, n_user, B), n_user + zipf_ids(rng, n_item, B)], 1).astype(np.int64)
        base = n_user + n_item
        cols = []
        for c in range(8):
            cols.append(base + rng.integers(0, per_ctx, B))
            base += per_ctx
        F1 = np.stack(cols, 1).astype(np.int64)
        Y = (n_user + rng.integers(0, n_item, (B, NG))).astype(np.int64)
        stg = Staging(torch.int32, dev)
        host, stride = pack_records([X, F1, Y], M, stg)
        recs.append(stg.upload(host.numel()).view(B, stride).clone())
        del stg
    m = OUR(8, 0, M, n_user, n_item, K, 0.1, 0.0, "AdagradOptimizer", True, False)     # lamda = 0: IndexedSlices rows
    ev = []

    def step(i):
        m.fit_device(recs[i % 3], 8, 0, NG)
    ms = timed(step, args.steps, 3)
    uniq = int(m._touch.count.item())                 # rows touched by the last step
    algo = 80 + 20 * 4 * K + 11 * 4 * K + 20 * K * uniq / B
    del ev
    return {"config": "c5-shaped HHFM (F=10, NG=10, M=10^7, K=128, sparse Adagrad rows), B=2^20, %d touched rows/step" % uniq,
            "ms_per_step": ms, "samples_per_s": B / ms * 1e3, "algorithmic_bytes_per_sample": algo,
            "roofline": {"bound": "hbm", "kernel": "pairrank_sum_train_staged_kernel + opt_rows_kernel (whole step)",
                         "achieved": B * algo / ms / 1e6, "peak": peaks(), "unit": "GB/s", "frac": B * algo / ms / 1e6 / peaks(),
                         "frac_of_nominal_8000_gbs": B * algo / ms / 1e6 / 8000.0,
                         "traffic": _traffic("pairrank_sum_train_staged_kernel")[0],
                         "traffic_capture": _traffic("pairrank_sum_train_staged_kernel")[1],
                         "traffic_note": "DRAM bytes per launch of the scatter kernel alone (16.7 GB algorithmic): the vector "
                                         "reductions into DRAM-resident gradient lines are read-modify-writes",
                         "peak_source": "measured (MEASURED_PEAKS.json)"}}


def run_fm_c5_l2(args, dev):
    """c5 scaled FM with the reference's regulariser (lamda = 0.1, FM.py:40,124): the dense L2 term moves all 10^7 rows every
    step in the reference; here it runs lazily (rows replay their missed `g = lamda*w` steps when they are next gathered,
    bit-identical to the dense update -- tests/test_gpu_fused_step.py).  Two roofline views: the bytes the lazy algorithm
    has to move, and the bytes of the reference's dense update it replaces (+ 20*M*K B per step)."""
    import torch
    from hhfm_b200.models import FM
    rng = np.random.default_rng(5)
    B, K = 1 << 20, 128
    n_user, n_item, n_ctx_ids = 4_000_000, 1_000_000, 5_000_000
    M = n_user + n_item + n_ctx_ids
    per_ctx = n_ctx_ids // 8
    batches = []
    for _ in range(3):
        cols = [zipf_ids(rng, n_user, B), n_user + zipf_ids(rng, n_item, B)]
        base = n_user + n_item
        for c in range(8):
            cols.append(base + rng.integers(0, per_ctx, B))
            base += per_ctx
        X = np.stack(cols, 1).astype(np.int32)
        batches.append((torch.from_numpy(X).to(dev), torch.from_numpy(rng.choice([1.0, 0.0], B).astype(np.float32)).to(dev)))
    m = FM(10, M, n_user, n_item, K, 0.1, 0.1, 1, "AdagradOptimizer", 0, 0)
    assert m._lazy()
    ms = timed(lambda i: m.fit_device(*batches[i % 3]), args.steps, 3)
    uniq = int(m._touch.count.item())
    F = 10
    kern = (4 * F + 4 * F * K + 4 * F + 4 + 4) + (4 * F * K + 4 * F)
    lazy = kern + (16 * K + 20 * K + 20) * uniq / B          # replay r/w (w, acc) + step r g, r/w w, r/w acc, w g=0 (+ bias)
    dense = kern + 20.0 * M * (K + 1) / B                     # what the reference's dense Adagrad moves per step
    t0 = __import__("time").perf_counter()
    m.flush(); torch.cuda.synchronize()
    flush_ms = (__import__("time").perf_counter() - t0) * 1e3
    return {"config": "c5 scaled FM with lamda = 0.1 (lazy-exact dense L2), B=2^20, %d gathered rows/step" % uniq,
            "ms_per_step": ms, "samples_per_s": B / ms * 1e3, "flush_ms": flush_ms,
            "algorithmic_bytes_per_sample": lazy, "reference_dense_bytes_per_sample": dense,
            "roofline": {"bound": "hbm", "achieved_gbs": B * lazy / ms / 1e6, "peak_gbs": peaks(), "frac": B * lazy / ms / 1e6 / peaks(),
                         "frac_vs_reference_dense_update": B * dense / ms / 1e6 / peaks(),
                         "note": "frac counts the bytes of the lazy algorithm; the reference's dense update of all 10^7 rows would "
                                 "need %.1f ms per step at the HBM peak on its own" % (20.0 * M * (K + 1) / peaks() / 1e6)}}


def run_bpr_c4(args, dev):
    import torch
    from hhfm_b200.models import BPR
    from hhfm_b200.engine import Staging, pack_records
    rng = np.random.default_rng(4)
    B, K, NG = 1 << 20, 128, 10
    n_user, n_item = 6522, 580
    M = 7730
    recs = []
    for _ in range(4):
        X = np.stack([zipf_ids(rng, n_user, B), n_user + zipf_ids(rng, n_item, B)], 1).astype(np.int64)
        Y = (n_user + rng.integers(0, n_item, (B, NG))).astype(np.int64)
        stg = Staging(torch.int32, dev)
        host, stride = pack_records([X, Y], M, stg)
        recs.append(stg.upload(host.numel()).view(B, stride).clone())
    m = BPR(M, n_user, n_item, K, 0.05, 0.01, "AdagradOptimizer")
    ms = timed(lambda i: m.fit_device(recs[i % 4], 0, 0, NG), args.steps, 5)
    algo = 12 * 4 * K + 48 + 3 * 4 * K
    return {"config": "c4 BPR restaurant shape (M=7730, N=580, K=128, NG=10, dense-L2 Adagrad), B=2^20",
            "ms_per_step": ms, "samples_per_s": B / ms * 1e3, "algorithmic_bytes_per_sample": algo,
            "roofline": l2_roofline(B, algo, ms, dev)}


def run_afm_c3(args, dev):
    import torch
    from hhfm_b200.models import AFM
    rng = np.random.default_rng(3)
    B, K = 1 << 17, 64
    batches = []
    for _ in range(2):
        X, M, n_user, n_item = frappe_rows(rng, B)
        batches.append((torch.from_numpy(X).to(dev), torch.from_numpy(rng.choice([1.0, -1.0], B).astype(np.float32)).to(dev)))
    m = AFM(n_user, n_item, M, 1, [K, K], "relu", 0.1, 100.0, [1, 1], "AdagradOptimizer", 0.999, 10)
    ms = timed(lambda i: m.fit_device(*batches[i % 2]), max(3, args.steps // 4), 2)
    P = 45
    flops = 3 * 2 * P * K * K + 10 * P * K
    ach = B * flops / ms / 1e9
    return {"config": "c3 AFM frappe-10 (P=45 pairs, K=A=64), B=2^17, fused tcgen05 training kernel (3xTF32)", "ms_per_step": ms,
            "samples_per_s": B / ms * 1e3, "algorithmic_flops_per_sample": flops,
            "roofline": {"bound": "tensor", "achieved_tflops": ach, "peak_tflops": tensor_peak_tf() / 6.0,
                         "frac": ach / (tensor_peak_tf() / 6.0), "frac_of_fp32_cuda_core_peak": ach / FP32_SIMT_TFLOPS,
                         "note": "the three products run on tcgen05; the step is bound by the SIMT phases between them (softmax, "
                                 "chain rule, operand construction), see profiles/r2_afm_tc_summary.md"}}


def run_dfm(args, dev):
    import torch
    from hhfm_b200.models import DeepFM
    rng = np.random.default_rng(6)
    B, K = 1 << 17, 64
    layers = [150, 200, 150]
    batches = []
    for _ in range(2):
        X, M, n_user, n_item = frappe_rows(rng, B)
        batches.append((torch.from_numpy(X).to(dev), torch.from_numpy(rng.choice([1.0, -1.0], B).astype(np.float32)).to(dev)))
    m = DeepFM(n_user, n_item, M, 10, K, layers, "relu", 0.01, 0, 0.01)
    ms = timed(lambda i: m.fit_device(*batches[i % 2]), max(3, args.steps // 2), 2)
    dims = [10 * K] + layers
    mm = sum(dims[i] * dims[i + 1] for i in range(3))
    flops = 3 * 2 * mm
    ach = B * flops / ms / 1e9
    ceiling = tensor_peak_tf() / 2.0 / 3.0            # tf32 runs at half the bf16 rate, three MMAs per fp32-grade product
    return {"config": "DeepFM frappe-10 (640-150-200-150, K=64), B=2^17", "ms_per_step": ms, "samples_per_s": B / ms * 1e3,
            "algorithmic_flops_per_sample": flops,
            "roofline": {"bound": "tensor", "achieved_tflops": ach, "peak_tflops": ceiling, "frac": ach / ceiling,
                         "frac_of_fp32_cuda_core_peak": ach / FP32_SIMT_TFLOPS,
                         "note": "GEMMs run as 3xTF32 splits on tcgen05 (fp32-grade accuracy); the ceiling of the scheme is the "
                                 "measured burst bf16 peak / 2 (tf32 rate) / 3 (three MMAs per product) in fp32-equivalent TFLOP/s; "
                                 "the whole step (split / transposition passes, scatter epilogue, optimizer) is timed, not the GEMMs alone"}}


RUNNERS = {"hhfm_c5": run_hhfm_c5, "fm_c1": run_fm_c1, "fm_c5": run_fm_c5, "fm_c5_l2": run_fm_c5_l2, "bpr_c4": run_bpr_c4, "afm_c3": run_afm_c3, "dfm": run_dfm}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=",".join(RUNNERS))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    out = {}
    for name in args.only.split(","):
        r = RUNNERS[name](args, dev)
        out[name] = r
        print(json.dumps({name: r}), flush=True)
        torch.cuda.empty_cache()
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
