"""Drop-in for the reference's Newcode/main.py (:50-63): the model sweep over the datasets.  Run it from a directory that
has `../data/positive/<dataset>/<dataset>.libfm` like the reference (or pass `--path` through the X_main argv)."""
from hhfm_b200.Newcode.AFM import AFM_main
from hhfm_b200.Newcode.BPR import BPR_main  # noqa: F401  (imported by the reference too)
from hhfm_b200.Newcode.CARS2 import CARS2_main
from hhfm_b200.Newcode.DFM import DFM_main
from hhfm_b200.Newcode.FM import FM_main
from hhfm_b200.Newcode.OurModel7 import M7_main  # HHFM


def run(datasets=('jiaju', 'resturant', 'frappe'), factors=(128,), argv=None):
    """main.py:50-63: TopK = 1 for jiaju, 5 otherwise; every model on every dataset."""
    sessions = []
    for f in factors:
        for data in datasets:
            topk = 1 if data == 'jiaju' else 5
            for fn in (M7_main, AFM_main, FM_main, DFM_main, CARS2_main):
                sessions.append(fn(data, f, topk, argv))
    return sessions


if __name__ == '__main__':
    run()
