"""Drop-in for the reference's Newcode/DFM.py: `parse_args`, `DeepFM`, `Train`, `DFM_main` (DFM.py:19-47,49-232,233-402,403)."""
import argparse

from hhfm_b200.models import DeepFM  # noqa: F401
from hhfm_b200.trainer import PointwiseTrain, default_result_file
from hhfm_b200.Newcode import NewLoadData as DATA

method = 'DFM'


def parse_args(dataname, factor, Topk, argv=None):
    """Same flags and defaults as DFM.py:19-47."""
    parser = argparse.ArgumentParser(description="Run .")
    parser.add_argument('--path', nargs='?', default='../data/positive/', help='Input data path.')
    parser.add_argument('--dataset', nargs='?', default=dataname, help='Choose a dataset.')
    parser.add_argument('--epoch', type=int, default=60, help='Number of epochs.')
    parser.add_argument('--batch_size', type=int, default=5000, help='Batch size.')
    parser.add_argument('--hidden_factor', type=int, default=factor, help='Number of hidden factors.')
    parser.add_argument('--lamda', type=float, default=0.01, help='Regularizer for bilinear part.')
    parser.add_argument('--keep', type=float, default=1)
    parser.add_argument('--lr', type=float, default=0.01, help='Learning rate.')
    parser.add_argument('--optimizer', nargs='?', default='AdagradOptimizer')
    parser.add_argument('--verbose', type=int, default=10)
    parser.add_argument('--batch_norm', type=int, default=0)
    parser.add_argument('--TopK', type=int, default=Topk)
    parser.add_argument('--Result', type=int, default=0, help='0:iteration 1:factors')
    return parser.parse_args(argv)


class Train(PointwiseTrain):
    method = method
    NG = 2
    neg_label = -1           # DFM.py:286
    topk_rows = 60           # DFM.py:367
    early_stop_cap = None    # DFM.py:299: no `or epoch>100`

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
            print("DFM: dataset=%s, factors=%d, #epoch=%d, batch=%d, lr=%.4f, lambda=%.1e, keep=%.2f, optimizer=%s, batch_norm=%d"
                  % (args.dataset, args.hidden_factor, args.epoch, args.batch_size, args.lr, args.lamda, args.keep,
                     args.optimizer, args.batch_norm))
        # DFM.py:257-259: layers [150,200,150], relu, l2_reg = args.lamda
        self.model = DeepFM(self.n_user, self.n_item, self.data.features_M, self.data.Train_data.shape[1] - 1,
                            args.hidden_factor, [150, 200, 150], 'relu', args.lr, args.verbose, args.lamda)


def DFM_main(dataname, factor, Topk, argv=None):
    args = parse_args(dataname, factor, Topk, argv)
    session = Train(args)
    session.train()
    return session
