"""Drop-in for the reference's Newcode/CARS2.py: `parse_args`, `CARS2`, `Train`, `CARS2_main` (CARS2.py:17-43,45-187,189-358,360)."""
import argparse

import numpy as np
import torch

from hhfm_b200 import engine
from hhfm_b200.models import CARS2  # noqa: F401
from hhfm_b200.trainer import PairwiseTrain, default_result_file, shuffle_rows
from hhfm_b200.Newcode import NewLoadData as DATA

method = 'CARS2'


def parse_args(dataname, factor, Topk, argv=None):
    """Same flags and defaults as CARS2.py:17-43."""
    parser = argparse.ArgumentParser(description="Run .")
    parser.add_argument('--path', nargs='?', default='../data/positive/')
    parser.add_argument('--dataset', nargs='?', default=dataname)
    parser.add_argument('--epoch', type=int, default=60)
    parser.add_argument('--batch_size', type=int, default=5000)
    parser.add_argument('--hidden_factor', type=int, default=factor)
    parser.add_argument('--lamda', type=float, default=0.001)
    parser.add_argument('--keep', type=float, default=1)
    parser.add_argument('--lr', type=float, default=0.01)
    parser.add_argument('--optimizer', nargs='?', default='AdagradOptimizer')
    parser.add_argument('--batch_norm', type=int, default=0)
    parser.add_argument('--TopK', type=int, default=Topk)
    parser.add_argument('--Result', type=int, default=0)
    return parser.parse_args(argv)


class Train(PairwiseTrain):
    method = method
    NG = 1                   # CARS2.py:248
    topk_rows = 300
    early_stop_after = 10    # CARS2.py:260

    def __init__(self, args):
        self.args = args
        self.batch_size = args.batch_size
        self.epoch = args.epoch
        self.TopK = args.TopK
        self.result_file = default_result_file()
        self.data = DATA.LoadData(self.args.path, self.args.dataset)
        self.n_user = self.data.n_user
        self.n_item = self.data.n_item
        # every distinct context tuple is one "feature" (CARS2.py:221-225; the reference numbers them in Python-set order,
        # here in lexicographic order -- the ids are arbitrary labels)
        ctx = np.asarray(self.data.Total_data.values[:, 3:], dtype=np.int64)
        self._ctx_keys = np.unique(ctx, axis=0)
        self.features_M = len(self._ctx_keys)
        self.feature_inject = {tuple(k): i for i, k in enumerate(self._ctx_keys.tolist())}
        print("OurModel: dataset=%s, factors=%d, #epoch=%d, batch=%d, lr=%.4f, lambda=%.1e, keep=%.2f, optimizer=%s, batch_norm=%d"
              % (args.dataset, args.hidden_factor, args.epoch, args.batch_size, args.lr, args.lamda, args.keep,
                 args.optimizer, args.batch_norm))
        self.model = CARS2(self.features_M, self.n_user, self.n_item, args.hidden_factor, args.lr, args.lamda,
                           args.optimizer)

    def context_ids(self, rows):
        """feature_inject[tuple(row[2:])] for id rows [n, F] (vectorised lexicographic search)."""
        q = np.ascontiguousarray(np.asarray(rows, dtype=np.int64)[:, 2:])
        dt = np.dtype([("f%d" % i, np.int64) for i in range(q.shape[1])])
        pos = np.searchsorted(self._ctx_keys.view(dt).reshape(-1), q.view(dt).reshape(-1))
        return np.clip(pos, 0, len(self._ctx_keys) - 1)

    def run_epoch(self):
        pos = self._train_values()[:, 1:]                    # persistent view: the in-place shuffle composes across epochs
        shuffle_rows(pos)                                    # CARS2.py:247 np.random.shuffle
        neg = self.sample_negative(pos, self.NG)
        fea = self.context_ids(pos)
        loss = 0
        for c0 in range(0, len(pos), self.batch_size):
            sl = slice(c0, c0 + self.batch_size)
            loss = loss + self.model.partial_fit({'X': np.array(pos[sl, :2], dtype=np.int64), 'Y': np.array(neg[sl], dtype=np.int64),
                                                  'F1': fea[sl]})
        return loss

    def score_rows(self, rows):
        return self.model.positive_feedback(rows[:, :2], self.context_ids(rows))

    def evaluate_AUC(self, data1):
        saved, self.device_sampler = self.device_sampler, False      # the device path scores id rows; CARS2 scores tuple ids
        try:
            return super().evaluate_AUC(data1)
        finally:
            self.device_sampler = saved

    def evaluate_TopK(self, data1):
        """CARS2.py:321-357."""
        dat = np.asarray(data1.values if hasattr(data1, "values") else data1)
        size = np.min([3000, len(dat)])
        codes = []
        for _ in range(int(size / self.topk_rows)):
            rows = np.array(dat[:, 1:][np.random.randint(0, len(dat), self.topk_rows)], dtype=np.int64)
            self.score = self.model.topk({'X': rows[:, 0], 'F1': self.context_ids(rows)}, 20)
            dev = self.model.device
            pred = torch.as_tensor(np.ascontiguousarray(self.score), dtype=torch.int32, device=dev) + self.n_user
            target = torch.as_tensor(rows[:, 1].astype(np.int32), device=dev)
            in_pf = torch.as_tensor(self.data.in_positive_feedback(rows).astype(np.uint8), device=dev)
            codes.append(engine.metrics_walk(pred.contiguous(), target, in_pf, self.TopK).cpu().numpy())
        codes = np.concatenate(codes) if codes else np.zeros(0, np.int32)
        return engine.metrics_from_codes(codes)


def CARS2_main(dataname, factor, Topk, argv=None):
    args = parse_args(dataname, factor, Topk, argv)
    session = Train(args)
    session.train()
    return session
