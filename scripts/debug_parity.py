"""GPU debug: where do post-step weights differ from the oracle (gradient vs optimizer)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import hhfm_oracle as O
from hhfm_b200.models import BPR, OUR

def report(name, got, ref, init, g=None, acc0=None):
    err = np.abs(got.astype(np.float64) - ref)
    disp = ref - init
    i = np.unravel_index(np.argmax(err), err.shape)
    print(name, 'max err %.3e at %s  ref %.6e got %.6e init %.6e disp_rms %.3e w_rms %.3e' % (err.max(), i, ref[i], got[i], init[i], np.sqrt((disp**2).mean()), np.sqrt((ref**2).mean())))
    if g is not None:
        print('   g(oracle) at worst = %.6e ; sensitivity lr*acc0/(acc0+g^2)^1.5' % g[i])
    for thr in (1e-7, 1e-6, 1e-5):
        print('   n(err>%g) = %d' % (thr, (err > thr).sum()))

rng = np.random.default_rng(3)
n_user, n_item = 6522, 580
M, K = n_user + n_item + 100, 128
bpr = BPR(M, n_user, n_item, K, 0.01, 0.1, 'AdagradOptimizer')
V0 = bpr.get_weights()["feature_embeddings"].copy(); V = V0.copy(); acc = np.full_like(V, 1e-8)
for step in range(3):
    X = np.stack([rng.integers(0, n_user, 5000), n_user + rng.integers(0, n_item, 5000)], axis=1)
    Y = n_user + rng.integers(0, n_item, (5000, 10))
    loss_ref, _, _, dV = O.pairrank_loss_grads(V, X, Y, None, None, (0, 0, 0), 0.1)
    Vp = V.copy()
    V, acc = O.adagrad_dense(V, acc, dV, 0.01)
    loss = bpr.partial_fit({"X": X, "Y": Y})
    got = bpr.get_weights()["feature_embeddings"]
    print('step', step, 'loss', loss, loss_ref)
    report('bpr', got, V, Vp, dV)
    # resync to isolate per-step error
    bpr.load_weights({"feature_embeddings": V})
    bpr._opt.state["feature_embeddings"][0].copy_(torch.as_tensor(acc).to(bpr.device))

rng = np.random.default_rng(2)
n_user, n_item, M, K, fc = 300, 500, 1000, 64, 8
m = OUR(fc, 0, M, n_user, n_item, K, 0.1, 0.01, 'AdagradOptimizer', True, False)
V0 = m.get_weights()["feature_embeddings"].copy(); V = V0.copy(); acc = np.full_like(V, 0.1)
for step in range(3):
    B = 5000
    X = np.stack([rng.integers(0, n_user, B), n_user + rng.integers(0, n_item, B)], axis=1)
    F1 = rng.integers(n_user + n_item, M, (B, fc)); Y = n_user + rng.integers(0, n_item, (B, 10))
    loss_ref, _, _, dV = O.pairrank_loss_grads(V, X, Y, F1, None, (0, 0, 0), 0.01)
    Vp = V.copy()
    V, acc = O.adagrad_dense(V, acc, dV, 0.1)
    loss = m.partial_fit({"X": X, "F1": F1, "Y": Y})
    got = m.get_weights()["feature_embeddings"]
    print('step', step, 'loss', loss, loss_ref, 'max|g|', np.abs(dV).max())
    report('hhfm', got, V, Vp, dV)
    m.load_weights({"feature_embeddings": V})
    m._opt.state["feature_embeddings"][0].copy_(torch.as_tensor(acc).to(m.device))
