"""Drop-in for the reference's Newcode/BPR.py: `parse_args`, `BPR`, `Train`, `BPR_main` (BPR.py:18-43,45-136,139-293,296)."""
import argparse

from hhfm_b200.models import BPR  # noqa: F401
from hhfm_b200.trainer import PairwiseTrain, default_result_file
from hhfm_b200.Newcode import NewLoadData as DATA

method = 'BPR'


def parse_args(dataname, factor, Topk, argv=None):
    """Same flags and defaults as BPR.py:18-43 (epoch 1110, lamda 0.1, lr 0.01, Result 1)."""
    parser = argparse.ArgumentParser(description="Run .")
    parser.add_argument('--path', nargs='?', default='../data/positive/')
    parser.add_argument('--dataset', nargs='?', default=dataname)
    parser.add_argument('--epoch', type=int, default=1110)
    parser.add_argument('--batch_size', type=int, default=5000)
    parser.add_argument('--hidden_factor', type=int, default=factor)
    parser.add_argument('--lamda', type=float, default=0.1)
    parser.add_argument('--keep', type=float, default=1)
    parser.add_argument('--lr', type=float, default=0.01)
    parser.add_argument('--optimizer', nargs='?', default='AdagradOptimizer')
    parser.add_argument('--batch_norm', type=int, default=0)
    parser.add_argument('--TopK', type=int, default=Topk)
    parser.add_argument('--Result', type=int, default=1)
    return parser.parse_args(argv)


class Train(PairwiseTrain):
    method = method
    NG = 10
    topk_rows = 100          # BPR.py:265
    early_stop_after = 30

    def __init__(self, args):
        self.args = args
        self.batch_size = args.batch_size
        self.epoch = args.epoch
        self.TopK = args.TopK
        self.result_file = default_result_file()
        self.data = DATA.LoadData(self.args.path, self.args.dataset)
        self.n_user = self.data.n_user
        self.n_item = self.data.n_item
        self.features_M = self.data.features_M
        self.valid_dimension = self.data.Train_data.shape[1] - 1
        print("OurModel: dataset=%s, factors=%d, #epoch=%d, batch=%d, lr=%.4f, lambda=%.1e, keep=%.2f, optimizer=%s, batch_norm=%d"
              % (args.dataset, args.hidden_factor, args.epoch, args.batch_size, args.lr, args.lamda, args.keep,
                 args.optimizer, args.batch_norm))
        self.model = BPR(self.features_M, self.n_user, self.n_item, args.hidden_factor, args.lr, args.lamda,
                         args.optimizer)

    def score_rows(self, rows):
        return self.model.positive_feedback(rows[:, :2])


def BPR_main(dataname, factor, Topk, argv=None):
    args = parse_args(dataname, factor, Topk, argv)
    session = Train(args)
    session.train()
    return session
