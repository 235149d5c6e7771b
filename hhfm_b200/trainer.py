"""Host-side trainer logic shared by the `Newcode/*.py` drop-ins: the epoch loop, negative sampling, batch
assembly, evaluate_AUC / evaluate_TopK and the result.txt log format of the reference `Train` classes
(e.g. Newcode/FM.py:199-359, Newcode/OurModel7.py:309-496, Newcode/BPR.py:139-293).

What changed versus the reference: the pure-Python double loops (sample_negative FM.py:284-294, the per-row
metric walk FM.py:336-357) are vectorised / moved to the device, and `toolz.partition_all` + fancy indexing
became plain slicing.  What did not change: the order and kind of `np.random` draws (so a seeded run consumes
the same random stream as the reference), batch contents (the pairwise trainers shuffle ONE persistent copy of the train
rows in place, as the reference shuffles its `.values` view), labels, chunk sizes and the log text.
"""
from __future__ import annotations

import copy
import os
from time import time

import numpy as np
import torch

from . import engine


def sample_negative(loader, n_user, n_item, data, num=10):
    """`Train.sample_negative` (FM.py:284-294): `num` uniform item draws per row, re-drawn while the item is in
    positive_feedback[key(row)].  Same random stream as the reference: one block `randint`, then one scalar
    `randint` per rejection in (row, column) order -- only the rejected cells are visited in Python."""
    data = np.asarray(data)
    samples = np.random.randint(n_user, n_user + n_item, size=(len(data), num))
    bad = loader.in_positive_feedback(data, samples)
    if bad.any():
        kc = [c - 1 for c in loader.key_cols]
        for i, j in zip(*np.nonzero(bad)):
            pf = loader.positive_feedback[tuple(data[i, kc].tolist())]
            neg = samples[i, j]
            while neg in pf:
                neg = np.random.randint(n_user, n_user + n_item)
            samples[i, j] = neg
    return samples


class BaseTrain(object):
    """Common trainer.  Subclasses set: method (log tag), model construction, batch assembly, score function."""

    method = "?"
    topk_rows = 300          # rows per evaluate_TopK round (FM.py:331; DFM 60, BPR 100)
    auc_first_chunk_only = False   # OurModel7.py:461 / BPR.py:258 `return` inside the chunk loop
    early_stop_after = 30    # FM.py:259 (OurModel7.py:390 uses 20)
    early_stop_tol = -0.0075  # FM.py:261 (AFM.py:330 uses -0.01)
    early_stop_cap = 100     # FM.py:262 `or epoch>100` (DFM has no cap)
    result_file = "../result.txt"
    # by default the epoch is device-resident -- the train ids are uploaded once per epoch, the negatives
    # are drawn by the device sampler (counter-based generator: statistically the reference's sampler, not numpy's stream),
    # the batches are row ranges of that buffer, the epoch loss is summed on the device, and evaluate_AUC runs on the device.
    # HHFM_DEVICE_SAMPLER=0 selects the host path, which consumes exactly the reference's numpy random stream
    # (pinned against the reference's own code in tests/test_host_logic.py).
    device_sampler = bool(int(os.environ.get("HHFM_DEVICE_SAMPLER", "1")))
    record_align = 1         # stride alignment of expanded rows (pair-ranking records need 4)

    # ---- to be provided by subclasses ----
    def score_rows(self, rows):
        """Model score [n,1] for id rows [n,F] (label column removed)."""
        raise NotImplementedError

    def run_epoch(self):
        """One epoch of minibatch training; returns the summed loss."""
        raise NotImplementedError

    # ---- shared logic ----
    def sample_negative(self, data, num=10):
        return sample_negative(self.data, self.n_user, self.n_item, data, num)

    def _sampler(self):
        s = getattr(self, "_dev_sampler", None)
        if s is None:
            s = self._dev_sampler = engine.DeviceSampler(self.data, self.n_user, self.n_item, self.model.device)
        return s

    def evaluate_AUC_device(self, X):
        """evaluate_AUC with everything after the id upload on the device: 50 negatives per row (device sampler), the
        negative rows built by replacing the item column, both sides scored by the forward kernel, wins counted."""
        if self.auc_first_chunk_only:
            X = X[:600]                                   # OurModel7.py:461 / BPR.py:258: `return` inside the chunk loop
        if len(X) == 0:
            return float("nan")
        smp, model = self._sampler(), self.model
        F = X.shape[1]
        ostride = (F + self.record_align - 1) // self.record_align * self.record_align
        wins = torch.zeros(1, dtype=torch.int64, device=model.device)
        step = max(600, (1 << 21) // 50)
        for c0 in range(0, len(X), step):
            rows = X[c0:c0 + step]
            idx, _ = model._uploader.upload([rows], model._M, align=self.record_align)
            negs = smp.sample(smp.key_ids(rows), 50)
            neg_rows = engine.expand_rows(idx, F, negs, ostride)
            engine.auc_wins(model.score_device(idx), model.score_device(neg_rows), 50, wins)
        return float(wins.item()) / (50.0 * len(X))

    def _train_values(self):
        """The train table as ONE persistent ndarray.  The reference takes `Train_data.values[:, 1:]` and shuffles it in place
        every epoch (OurModel7.py:369-370): with the pandas of its time that is a view, so the permutation persists, composes
        across epochs, and `evaluate_AUC(Train_data)` (first 600 rows only, :461) sees the shuffled rows.  pandas 3 hands out
        read-only copies, so the drop-in keeps the array itself."""
        tv = getattr(self, "_train_tv", None)
        if tv is None:
            tv = self._train_tv = np.array(self.data.Train_data.values)
        return tv

    def evaluate_AUC(self, data1):
        """FM.py:296-324: 50 sampled negatives per positive, fraction with pos > neg, chunks of 600 rows."""
        if data1 is self.data.Train_data and getattr(self, "_train_tv", None) is not None:
            data1 = self._train_tv
        dat = np.asarray(data1.values if hasattr(data1, "values") else data1)
        dat = dat[dat[:, 0] > 0]
        X = np.array(dat[:, 1:], dtype=np.int64)
        if self.device_sampler:
            return self.evaluate_AUC_device(X)
        score = []
        for c0 in range(0, len(X), 600):
            pos = X[c0:c0 + 600]
            negs = self.sample_negative(pos, 50)
            neg = np.tile(pos[:, None, :], [1, 50, 1]).reshape(-1, pos.shape[1])
            neg[:, 1] = negs.reshape(-1)
            neg_score = self.score_rows(neg)
            pos_score = self.score_rows(pos)
            pos_rep = np.reshape(np.tile(np.expand_dims(pos_score, axis=1), [1, 50, 1]), [-1, 1])
            score.extend(np.reshape(pos_rep > neg_score, [-1]).tolist())
            if self.auc_first_chunk_only:
                return np.mean(score)
        return np.mean(score)

    def evaluate_TopK(self, data1):
        """FM.py:325-359: int(size/num) rounds of `num` rows drawn with replacement, top-20 retrieval on the device,
        then the HR / NDCG / reciprocal-rank walk (device kernel, including the positive_feedback quirk)."""
        dat = np.asarray(data1.values if hasattr(data1, "values") else data1)
        size = np.min([3000, len(dat)])
        num = self.topk_rows
        codes = []
        for _ in range(int(size / num)):
            rows = np.array(dat[:, 1:][np.random.randint(0, len(dat), num)], dtype=np.int64)
            self.score = self.model.topk(rows, 20)
            pred = torch.as_tensor(np.ascontiguousarray(self.score), dtype=torch.int32, device=self.model.device) + self.n_user
            target = torch.as_tensor(rows[:, 1].astype(np.int32), device=self.model.device)
            in_pf = torch.as_tensor(self.data.in_positive_feedback(rows).astype(np.uint8), device=self.model.device)
            codes.append(engine.metrics_walk(pred.contiguous(), target, in_pf, self.TopK).cpu().numpy())
        codes = np.concatenate(codes) if codes else np.zeros(0, np.int32)
        return engine.metrics_from_codes(codes)

    def _log(self, text):
        print(text)
        try:
            with open(self.result_file, "a") as f:
                f.write(text + "\n")
        except OSError:
            pass

    def _eval_line(self, head, t_train, t2):
        a_tr = self.evaluate_AUC(self.data.Train_data)
        a_te = self.evaluate_AUC(self.data.Test_data)
        tk = self.evaluate_TopK(self.data.Test_data)
        if t_train is None:
            return "%s Init: \t train=AUC:%.4f;test=AUC:%.4f,HR:%.4f,NDCG:%.4f,PRE:%.4f;[%.1f s]" % (
                head, a_tr, a_te, tk[0], tk[1], tk[2], time() - t2)
        return "%s [%.1f s]\ttrain=AUC:%.4f;test=AUC:%.4f,HR:%.4f,NDCG:%.4f,PRE:%.4f;[%.1f s]" % (
            head, t_train, a_tr, a_te, tk[0], tk[1], tk[2], time() - t2)

    def train(self):
        """The epoch loop of FM.py:221-282: optional initial evaluation, `range(1, epoch)` epochs, evaluation
        every `verbose` epochs (Result == 0) or loss-plateau early stop (Result == 1)."""
        args = self.args
        t2 = time()
        if args.Result == 0:
            self._log(self._eval_line("Dataset=%s %s" % (args.dataset, self.method), None, t2))
        self.loss_epoch = []
        every = getattr(self, "verbose", 10)      # FM.py:273 `self.verbose > 0 and epoch % self.verbose == 0`; OurModel7.py:403 `% 10`
        for epoch in range(1, self.epoch):
            t1 = time()
            loss = self.run_epoch()
            self.loss_epoch.append(loss)
            t2 = time()
            if args.Result == 1 and epoch > self.early_stop_after:
                n = 3
                le = np.array(self.loss_epoch)
                condition = np.sum((le[-1 - n:-1] / le[-2 - n:-2] - 1) > self.early_stop_tol)
                if condition == n or (self.early_stop_cap is not None and epoch > self.early_stop_cap):
                    self._log(self._eval_line("%s%s Epoch %d" % (args.dataset, self.method, epoch), t2 - t1, t2))
                    break
            if args.Result == 0 and every > 0 and epoch % every == 0:
                self._log(self._eval_line("%s Epoch %d" % (self.method, epoch), t2 - t1, t2))


class PointwiseTrain(BaseTrain):
    """FM / AFM / DFM style epochs (FM.py:240-256): positives + NG sampled-item copies, shuffled, chunked."""

    NG = 2
    neg_label = 0            # FM.py:248 `-0`; AFM.py:317 / DFM.py:286 use -1
    n_model_cols = None      # id columns the model reads (None = all; MF reads user and item only, MF.py:81-82)

    def run_epoch_device(self):
        """The same epoch with the negatives drawn on the device and the shuffled batches cut from device-resident rows:
        one id upload per epoch instead of one per batch."""
        model, smp = self.model, self._sampler()
        pos = np.array(self.data.Train_data.values)
        n = pos.shape[0]
        F = pos.shape[1] - 1 if self.n_model_cols is None else int(self.n_model_cols)
        idx, _ = model._uploader.upload([pos[:, 1:1 + F]], model._M, align=1)
        kid = getattr(self, "_train_kid_dev", None)          # the train table is the same every epoch: look its keys up once
        if kid is None or kid.numel() != n:
            kid = self._train_kid_dev = smp.key_ids(pos[:, 1:])
        negs = smp.sample(kid, self.NG)
        rows = torch.cat([idx, engine.expand_rows(idx, F, negs)], dim=0)
        y = torch.cat([torch.as_tensor(pos[:, 0].astype(np.float32), device=model.device),
                       torch.full((n * self.NG,), float(self.neg_label), device=model.device)])
        perm = torch.as_tensor(np.random.permutation(len(rows)), device=model.device)      # FM.py:250 np.random.shuffle
        rows, y = rows[perm].contiguous(), y[perm].contiguous()
        loss = torch.zeros(1, dtype=torch.float64, device=model.device)     # one read-back per epoch instead of one per batch
        for c0 in range(0, len(rows), self.batch_size):
            model.fit_device(rows[c0:c0 + self.batch_size], y[c0:c0 + self.batch_size])
            loss += model._loss_dev.double()
        return float(loss.item())

    def run_epoch(self):
        if self.device_sampler:
            return self.run_epoch_device()
        pos = np.array(self.data.Train_data.values)
        neg = np.tile(np.expand_dims(copy.deepcopy(pos), axis=1), [1, self.NG, 1]).reshape(-1, pos.shape[1])
        neg[:, 2] = self.sample_negative(pos[:, 1:], self.NG).reshape(-1)
        neg[:, 0] = self.neg_label
        dat = np.append(pos, neg, axis=0)
        shuffle_rows(dat)                                    # FM.py:250 np.random.shuffle
        fit = _PipelinedFit(self.model)
        for c0 in range(0, len(dat), self.batch_size):
            chunk = dat[c0:c0 + self.batch_size]
            fit({'X': np.array(chunk[:, 1:], dtype=np.int64), 'Y': np.expand_dims(chunk[:, 0], axis=1)})
        return fit.total()

    def score_rows(self, rows):
        return self.model.predict(rows)


class PairwiseTrain(BaseTrain):
    """OurModel7 / BPR style epochs (OurModel7.py:369-387): shuffled positives with NG=10 sampled negatives."""

    NG = 10
    auc_first_chunk_only = True
    context = False
    time = False
    time_dimension = 0

    def split(self, rows):
        """Batch dict for `partial_fit` from id rows [n,F] (OurModel7.py:374-385)."""
        d = {'X': np.array(rows[:, :2], dtype=np.int64)}
        if self.context and self.time:
            d['F1'] = np.array(rows[:, 2:-self.time_dimension], dtype=np.int64)
            d['F2'] = np.array(rows[:, -self.time_dimension:], dtype=np.int64)
        elif self.context:
            d['F1'] = np.array(rows[:, 2:], dtype=np.int64)
        elif self.time:
            d['F2'] = np.array(rows[:, 2:], dtype=np.int64)
        return d

    record_align = 4

    def run_epoch_device(self):
        """Shuffled positives are uploaded once per epoch as records with NG empty negative slots, the device sampler
        fills the slots, and the batches are row ranges of that buffer."""
        model, smp = self.model, self._sampler()
        pos = self._train_values()[:, 1:]                    # persistent view: the shuffle composes across epochs
        kid = getattr(self, "_train_kid", None)              # key id of every row, carried through the shuffles
        if kid is None or len(kid) != len(pos):
            kid = self.data.key_ids(pos).astype(np.int32)
        perm = shuffle_rows(pos)                             # OurModel7.py:370 np.random.shuffle
        kid = self._train_kid = kid[perm]
        d = self.split(pos)
        parts = [d['X']] + ([d['F1']] if 'F1' in d else []) + ([d['F2']] if 'F2' in d else [])
        width = sum(p.shape[1] for p in parts)
        rec, stride = model._uploader.upload(parts, model._M, align=4, extra_cols=self.NG)
        smp.sample(torch.as_tensor(kid).to(model.device), self.NG, out=rec, out_stride=stride, out_col0=width)
        n_ctx = d['F1'].shape[1] if 'F1' in d else 0
        n_time = d['F2'].shape[1] if 'F2' in d else 0
        loss = torch.zeros(1, dtype=torch.float64, device=model.device)     # one read-back per epoch instead of one per batch
        for c0 in range(0, len(pos), self.batch_size):
            model.fit_device(rec[c0:c0 + self.batch_size], n_ctx, n_time, self.NG)
            loss += model._loss_dev.double()
        return float(loss.item())

    def run_epoch(self):
        if self.device_sampler:
            return self.run_epoch_device()
        pos = self._train_values()[:, 1:]                    # persistent view (see _train_values)
        perm = shuffle_rows(pos)                             # OurModel7.py:370 np.random.shuffle
        if getattr(self, "_train_kid", None) is not None:
            self._train_kid = self._train_kid[perm]          # the device path's cached key ids follow the rows
        neg = self.sample_negative(pos, self.NG)
        fit = _PipelinedFit(self.model)
        for c0 in range(0, len(pos), self.batch_size):
            d = self.split(pos[c0:c0 + self.batch_size])
            d['Y'] = np.array(neg[c0:c0 + self.batch_size], dtype=np.int64)
            fit(d)
        return fit.total()


def shuffle_rows(arr):
    """`np.random.shuffle(arr)` for a 2-D array, bit for bit -- same permutation, same state of numpy's global generator
    afterwards (both run the same Fisher-Yates draws; tests/test_host_logic.py) -- but as ONE index permutation and one
    gather instead of numpy's buffered row-by-row swaps, which cost 40 ms per epoch on the 88 571 frappe rows."""
    perm = np.random.permutation(len(arr))
    arr[:] = arr[perm]
    return perm


class _PipelinedFit:
    """The epoch loop's `loss = loss + model.partial_fit(batch)` (FM.py:251-256) with one step in flight: the loss of batch i
    is waited for after batch i+1 has been enqueued (`partial_fit_async`), so the host assembles and packs the next batch
    while the GPU runs the current one.  The losses are added in batch order; a model without the async call (a test
    double) is driven synchronously."""

    def __init__(self, model):
        self._async = getattr(model, "partial_fit_async", None)
        self._sync = model.partial_fit
        self._prev = None
        self._loss = 0

    def __call__(self, batch):
        if self._async is None:
            self._loss = self._loss + self._sync(batch)
            return
        h = self._async(batch)
        if self._prev is not None:
            self._loss = self._loss + self._prev.result()
        self._prev = h

    def total(self):
        if self._prev is not None:
            self._loss = self._loss + self._prev.result()
            self._prev = None
        return self._loss


def default_result_file():
    return os.environ.get("HHFM_RESULT_FILE", "../result.txt")
