import os, sys, time, cProfile, pstats
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
import reference_bands as rb
from hhfm_b200.Newcode import OurModel7 as M7
os.environ.setdefault("HHFM_RESULT_FILE", os.devnull)
np.random.seed(1)
args = M7.parse_args("frappe", 64, 5, ["--path", rb.DATA, "--epoch", "2", "--Result", "2"])
sess = M7.Train(args)
sess.run_epoch(); sess.run_epoch()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
t0 = time.perf_counter()
for _ in range(3):
    sess.run_epoch()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
pr.disable()
print("epoch s", dt)
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
