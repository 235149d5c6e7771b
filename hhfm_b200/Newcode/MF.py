"""Drop-in for the reference's Newcode/MF.py: `parse_args`, `MF`, `Train` (MF.py:17-41,43-149,150-275).  The reference
file is a script (it parses `sys.argv`, reads a module-level `args` inside `Train.train` :181 and ends in a call to a
method that does not exist :295); here `parse_args` takes an optional argv, `Train` uses its own args, and `MF_main`
mirrors the `X_main(dataname, factor, TopK)` entries of the other models."""
import argparse
from time import time

import numpy as np

from hhfm_b200.models import MF  # noqa: F401
from hhfm_b200.trainer import PointwiseTrain, default_result_file
from hhfm_b200.Newcode import NewLoadData as DATA

method = 'MF'


def parse_args(argv=None, dataname='fra', factor=256):
    """Same flags and defaults as MF.py:17-41."""
    parser = argparse.ArgumentParser(description="Run FM.")
    parser.add_argument('--path', nargs='?', default='../data/positive/', help='Input data path.')
    parser.add_argument('--dataset', nargs='?', default=dataname, help='Choose a dataset.')
    parser.add_argument('--epoch', type=int, default=110, help='Number of epochs.')
    parser.add_argument('--batch_size', type=int, default=4096, help='Batch size.')
    parser.add_argument('--hidden_factor', type=int, default=factor, help='Number of hidden factors.')
    parser.add_argument('--lamda', type=float, default=0.01, help='Regularizer for bilinear part.')
    parser.add_argument('--keep', type=float, default=0.7,
                        help='Keep probility (1-dropout) for the bilinear interaction layer. 1: no dropout')
    parser.add_argument('--lr', type=float, default=0.01, help='Learning rate.')
    parser.add_argument('--optimizer', nargs='?', default='AdagradOptimizer')
    parser.add_argument('--verbose', type=int, default=5)
    parser.add_argument('--batch_norm', type=int, default=0)
    return parser.parse_args(argv)


class Train(PointwiseTrain):
    """MF.py:150-275: NG = 2 negatives labelled -1, batches of 4096, evaluation every `verbose` epochs; evaluate_AUC
    draws 10 negatives per positive in chunks of 10 000 (:214-239), evaluate_TopK is 10 rounds of 100 rows against the
    top-100 list with the plain `item in prediction` test (:240-258)."""
    method = method
    NG = 2
    neg_label = -1           # MF.py:190
    n_model_cols = 2         # the table holds users and items only (MF.py:171)

    def __init__(self, args):
        self.args = args
        self.batch_size = args.batch_size
        self.epoch = args.epoch
        self.verbose = args.verbose
        self.keep = args.keep
        self.TopK = 100
        self.result_file = default_result_file()
        self.data = DATA.LoadData(self.args.path, self.args.dataset)
        self.n_user = self.data.n_user
        self.n_item = self.data.n_item
        if args.verbose > 0:
            print("FM: dataset=%s, factors=%d, #epoch=%d, batch=%d, lr=%.4f, lambda=%.1e, keep=%.2f, optimizer=%s, batch_norm=%d"
                  % (args.dataset, args.hidden_factor, args.epoch, args.batch_size, args.lr, args.lamda, args.keep,
                     args.optimizer, args.batch_norm))
        # MF.py:171: the table holds the users and the items only
        self.model = MF(self.n_user + self.n_item, self.n_user, self.n_item, args.hidden_factor, args.lr, args.lamda,
                        args.keep, args.optimizer, args.batch_norm, args.verbose)

    def evaluate_AUC(self, data1):
        dat = np.asarray(data1.values if hasattr(data1, "values") else data1)
        dat = dat[dat[:, 0] > 0]
        X = np.array(dat[:, 1:], dtype=np.int64)
        score = []
        for c0 in range(0, len(X), 10000):
            pos = X[c0:c0 + 10000]
            negs = self.sample_negative(pos)                                  # 10 per row (MF.py:224)
            neg = np.tile(pos[:, None, :], [1, 10, 1]).reshape(-1, pos.shape[1])
            neg[:, 1] = negs.reshape(-1)
            neg_score = self.score_rows(neg)
            pos_score = np.reshape(np.tile(np.expand_dims(self.score_rows(pos), axis=1), [1, 10, 1]), [-1, 1])
            score.extend(np.reshape(pos_score > neg_score, [-1]).tolist())
        return np.mean(score)

    def evaluate_TopK(self, data1):
        dat = np.asarray(data1.values if hasattr(data1, "values") else data1)
        res_map, res_ndcg, res_pre = [], [], []
        for _ in range(10):                                                    # int(1000 / 100) rounds, MF.py:247
            rows = np.array(dat[:, 1:][np.random.randint(0, len(dat), 100)], dtype=np.int64)
            self.score = self.model.topk(rows)
            prediction = self.score + self.n_user
            for i, item in enumerate(rows[:, 1]):
                hit = np.flatnonzero(prediction[i] == item)
                if hit.size:
                    n = int(hit[0])
                    res_map.append(1); res_ndcg.append(np.log(2) / np.log(n + 2)); res_pre.append(1 / (n + 1))
                else:
                    res_map.append(0); res_ndcg.append(0); res_pre.append(0)
        return [np.average(res_map), np.average(res_ndcg), np.average(res_pre)]

    def train(self):
        """MF.py:173-212: initial evaluation, `range(1, epoch)` epochs, evaluation every `verbose` epochs (printed only --
        the reference's MF does not append to result.txt)."""
        t2 = time()
        a_tr, a_te, tk = self.evaluate_AUC(self.data.Train_data), self.evaluate_AUC(self.data.Test_data), \
            self.evaluate_TopK(self.data.Test_data)
        if self.verbose > 0:
            print("Init: \t train=AUC:%.4f;test=AUC:%.4f,HR:%.4f,NDCG:%.4f,PRE:%.4f;[%.1f s]," % (a_tr, a_te, tk[0], tk[1], tk[2], time() - t2))
        self.loss_epoch = []
        for epoch in range(1, self.epoch):
            t1 = time()
            self.loss_epoch.append(self.run_epoch())
            t2 = time()
            if self.verbose > 0 and epoch % self.verbose == 0:
                a_tr, a_te, tk = self.evaluate_AUC(self.data.Train_data), self.evaluate_AUC(self.data.Test_data), \
                    self.evaluate_TopK(self.data.Test_data)
                print("Epoch %d [%.1f s]\ttrain=AUC:%.4f;test=AUC:%.4f,HR:%.4f,NDCG:%.4f,PRE:%.4f;[%.1f s]"
                      % (epoch, t2 - t1, a_tr, a_te, tk[0], tk[1], tk[2], time() - t2))


def MF_main(dataname='fra', factor=256, Topk=100, argv=None):
    args = parse_args(argv, dataname, factor)
    session = Train(args)
    session.train()
    return session


if __name__ == '__main__':
    MF_main()
