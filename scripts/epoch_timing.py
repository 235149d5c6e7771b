#!/usr/bin/env python
"""Reference-style end-to-end epoch and evaluation time of the drop-in trainers ("additionally report the
reference-style end-to-end epoch (sampler included) for context against result.txt").

A frappe-sized synthetic libfm file (96 203 rows, 957 users, 4 082 items, the shipped file's 3 context columns) is
written to a temp directory; `FM_main` / `M7_main` run with the reference defaults (batch 5 000, K = 64) for a few
epochs, once with the stream-compatible host sampler and once with the device sampler (HHFM_DEVICE_SAMPLER).
result.txt of the reference (TF, CPU): FM frappe epoch 2-5 s, evaluation 26-40 s; HHFM epoch 2-3 s.
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_dataset(root, rows=96203, n_user=957, n_item=4082, seed=1):
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "frappe"), exist_ok=True)
    p = 1.0 / np.arange(1, n_user + 1) ** 1.1; p /= p.sum()
    u = rng.choice(n_user, rows, p=p)
    q = 1.0 / np.arange(1, n_item + 1) ** 1.1; q /= q.sum()
    it = rng.choice(n_item, rows, p=q)
    u[:n_user] = np.arange(n_user); it[:n_item] = np.arange(n_item)          # every token appears
    day = rng.integers(0, 7, rows); wk = rng.integers(0, 2, rows); hw = rng.integers(0, 3, rows)
    with open(os.path.join(root, "frappe", "frappe.libfm"), "w") as f:
        f.write("".join("1 u%d i%d d%d w%d h%d\n" % t for t in zip(u, it, day, wk, hw)))
    return os.path.join(root, "")


def main():
    from hhfm_b200 import trainer
    from hhfm_b200.Newcode import FM as FMmod, OurModel7 as M7mod
    tmp = tempfile.mkdtemp()
    path = write_dataset(tmp)
    os.environ["HHFM_RESULT_FILE"] = os.path.join(tmp, "result.txt")
    out = {}
    for name, mod, argv in (("FM", FMmod, ["--verbose", "0"]), ("M7", M7mod, [])):
        for dev_sampler in (False, True):
            trainer.BaseTrain.device_sampler = dev_sampler
            np.random.seed(1)
            args = mod.parse_args("frappe", 64, 10, ["--path", path, "--epoch", "2", "--Result", "1"] + argv)
            sess = mod.Train(args)
            sess.run_epoch()                                          # warm-up (hot-row plan, allocations)
            t0 = time.perf_counter(); n_ep = 3
            for _ in range(n_ep):
                sess.run_epoch()
            t_epoch = (time.perf_counter() - t0) / n_ep
            t0 = time.perf_counter()
            auc = sess.evaluate_AUC(sess.data.Train_data)
            t_auc = time.perf_counter() - t0
            t0 = time.perf_counter()
            tk = sess.evaluate_TopK(sess.data.Test_data)
            t_topk = time.perf_counter() - t0
            out["%s_%s_sampler" % (name, "device" if dev_sampler else "host")] = {
                "epoch_s": round(t_epoch, 4), "evaluate_AUC_train_s": round(t_auc, 4), "evaluate_TopK_s": round(t_topk, 4),
                "train_AUC": round(float(auc), 4), "HR": round(float(tk[0]), 4), "train_rows": int(len(sess.data.Train_data))}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
