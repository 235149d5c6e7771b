"""Drop-in for the reference's Newcode/WDMF.py: `parse_args`, `WD`, `Train`, `WDMF_main` (WDMF.py:22-49,51-126,132-257,265).

The reference's `WD` is tf.contrib.learn's canned DNNLinearCombinedClassifier; `hhfm_b200.models.WD` restates it on the
sm_100a kernels (parity unpinned: see the class docstring)."""
import argparse
from time import time

import numpy as np

from hhfm_b200.models import WD  # noqa: F401
from hhfm_b200.trainer import BaseTrain, default_result_file, shuffle_rows
from hhfm_b200.Newcode import NewLoadData as DATA

method = 'WD'


def parse_args(dataname, factor, Topk, argv=None):
    """Same flags and defaults as WDMF.py:22-49."""
    parser = argparse.ArgumentParser(description="Run .")
    parser.add_argument('--path', nargs='?', default='../data/positive/', help='Input data path.')
    parser.add_argument('--dataset', nargs='?', default=dataname, help='Choose a dataset.')
    parser.add_argument('--process', nargs='?', default='train', help='Process type: train, evaluate.')
    parser.add_argument('--mla', type=int, default=0)
    parser.add_argument('--epoch', type=int, default=110, help='Number of epochs.')
    parser.add_argument('--batch_size', type=int, default=4096, help='Batch size.')
    parser.add_argument('--hidden_factor', type=int, default=factor, help='Number of hidden factors.')
    parser.add_argument('--lamda', type=float, default=0.1, help='Regularizer for bilinear part.')
    parser.add_argument('--keep', type=float, default=1)
    parser.add_argument('--lr', type=float, default=0.01, help='Learning rate.')
    parser.add_argument('--optimizer', nargs='?', default='AdagradOptimizer')
    parser.add_argument('--verbose', type=int, default=10)
    parser.add_argument('--batch_norm', type=int, default=0)
    parser.add_argument('--TopK', type=int, default=Topk)
    # not in the reference: sizes of the restated estimator, so that tests and small boxes can shrink it
    parser.add_argument('--wd_steps', type=int, default=500, help='steps per partial_fit (WDMF.py:115 fit(steps=500))')
    parser.add_argument('--wd_outer', type=int, default=10, help='outer epochs (WDMF.py:167 range(1, 11))')
    parser.add_argument('--wd_hidden', nargs='?', default='1024,512,256')
    parser.add_argument('--wd_dim', type=int, default=128)
    return parser.parse_args(argv)


class Train(BaseTrain):
    """WDMF.py:132-257: ten outer epochs of (one negative per positive, shuffle, `partial_fit` = 500 full-batch steps), each
    followed by AUC on 10 sampled negatives per row and the `item in prediction` form of the top-K walk (vectorised here)."""
    method = method

    def __init__(self, args):
        self.args = args
        self.batch_size = args.batch_size
        self.epoch = args.epoch
        self.verbose = args.verbose
        self.keep = args.keep
        self.TopK = args.TopK
        self.result_file = default_result_file()
        self.data = DATA.LoadData(self.args.path, self.args.dataset)
        self.n_user = self.data.n_user
        self.n_item = self.data.n_item
        if args.verbose > 0:
            print("FM: dataset=%s, factors=%d, #epoch=%d, batch=%d, lr=%.4f, lambda=%.1e, keep=%.2f, optimizer=%s, batch_norm=%d"
                  % (args.dataset, args.hidden_factor, args.epoch, args.batch_size, args.lr, args.lamda, args.keep,
                     args.optimizer, args.batch_norm))
        hidden = [int(h) for h in str(args.wd_hidden).split(',')]
        self.model = WD(self.data.Train_data.shape[1] - 1, self.n_user, self.n_item,
                        features_M=max(100000, int(self.data.features_M)), hidden_units=hidden, embedding_dim=args.wd_dim,
                        steps=args.wd_steps)

    def score_rows(self, rows):
        return self.model.predict(rows)[:, 1].reshape(-1, 1)

    OUTER_NEGATIVES = 1        # WDMF.py:169  NG = 1
    AUC_NEGATIVES = 10         # WDMF.py:191,213  sample_negative(..., num=10)
    AUC_CHUNK = 3000           # WDMF.py:209
    TOPK_ROUNDS, TOPK_ROWS = 20, 50   # WDMF.py:227,234: int(500/25) rounds of 50 rows

    def _epoch_batch(self):
        """Positives plus one sampled-item copy each (label 0), shuffled: the (X, Y) of WDMF.py:168-184."""
        pos = np.asarray(self.data.Train_data.values)
        neg = np.repeat(pos, self.OUTER_NEGATIVES, axis=0)
        neg[:, 2] = self.sample_negative(pos[:, 1:], self.OUTER_NEGATIVES).ravel()
        neg[:, 0] = 0
        rows = np.concatenate([pos, neg], axis=0)
        shuffle_rows(rows)                                   # np.random.shuffle, same permutation and generator state
        return rows[:, 1:].astype(np.int64), rows[:, 0]

    def train(self):
        t0 = time()
        # the reference logs an all-zero "Init" line (its initial evaluation is commented out, WDMF.py:155-165)
        self._log("Dataset=%s %s Init: \t train=AUC:%.4f;test=AUC:%.4f,HR:%.4f,NDCG:%.4f,PRE:%.4f;[%.1f s]"
                  % (self.args.dataset, method, 0, 0, 0, 0, 0, time() - t0))
        self.loss_epoch = []
        for outer in range(1, int(self.args.wd_outer) + 1):
            t_fit = time()
            X, Y = self._epoch_batch()
            self.loss_epoch.append(self.model.partial_fit(X, Y))
            t_eval = time()
            auc_train = self.evaluate_AUC(self.data.Train_data)
            auc_test = self.evaluate_AUC(self.data.Test_data)
            hr, ndcg, pre = self.evaluate_TopK(self.data.Test_data)
            self._log("%s Epoch %d [%.1f s]\ttrain=AUC:%.4f;test=AUC:%.4f,HR:%.4f,NDCG:%.4f,PRE:%.4f;[%.1f s]"
                      % (method, outer * 10, t_eval - t_fit, auc_train, auc_test, hr, ndcg, pre, time() - t_eval))

    def evaluate_AUC(self, data1):
        """WDMF.py:202-225.  Quirk kept: `predict_proba(...)[:, 1]` is 1-D there, so `pos_score > neg_score` broadcasts a
        [10n, 1] column against a [10n] row -- every positive of a 3000-row chunk is compared with EVERY sampled negative of
        the chunk (10 per positive), not only with its own.  The mean of that [10n, 10n] matrix is computed here from the
        sorted negatives (same value, no 900 MB boolean matrix); the chunk means are averaged."""
        table = np.asarray(data1.values if hasattr(data1, "values") else data1)
        X = table[table[:, 0] > 0][:, 1:].astype(np.int64)
        shares = []
        for start in range(0, len(X), self.AUC_CHUNK):
            pos = X[start:start + self.AUC_CHUNK]
            neg = np.repeat(pos, self.AUC_NEGATIVES, axis=0)
            neg[:, 1] = self.sample_negative(pos).ravel()
            p_neg = np.sort(self.score_rows(neg).ravel())
            p_pos = self.score_rows(pos).ravel()
            below = np.searchsorted(p_neg, p_pos, side="left")          # negatives strictly below each positive
            shares.append(float(below.sum()) / (len(p_pos) * len(p_neg)))
        return np.mean(shares)

    def evaluate_TopK(self, data1):
        """WDMF.py:226-249: 20 rounds of 50 rows drawn with replacement; a row scores when its own item is in its top-K list
        (hit, log 2 / log(rank + 2), 1 / (rank + 1) with rank = position of the first match)."""
        table = np.asarray(data1.values if hasattr(data1, "values") else data1)
        hits, gains, recip = [], [], []
        for _ in range(self.TOPK_ROUNDS):
            feed = table[np.random.randint(0, len(table), self.TOPK_ROWS)][:, 1:].astype(np.int64)
            self.score = self.model.topk(feed, self.TopK)
            match = (self.score + self.n_user) == feed[:, 1:2]
            found = match.any(axis=1)
            rank = match.argmax(axis=1)
            hits.append(found.astype(np.float64))
            gains.append(np.where(found, np.log(2) / np.log(rank + 2.0), 0.0))
            recip.append(np.where(found, 1.0 / (rank + 1.0), 0.0))
        return [np.average(np.concatenate(hits)), np.average(np.concatenate(gains)), np.average(np.concatenate(recip))]


def WDMF_main(dataname, factor, Topk, argv=None):
    args = parse_args(dataname, factor, Topk, argv)
    session = Train(args)
    session.train()
    return session
