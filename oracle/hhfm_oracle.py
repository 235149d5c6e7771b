"""CPU oracle for the HHFM factorization-machine hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a NumPy fp32 restatement of the TensorFlow-1.x graphs the reference builds in
`Newcode/{FM,MF,AFM,DFM,OurModel7,BPR,CARS2}.py`, of the Wide&Deep estimator `Newcode/WDMF.py` wraps (TF's
documented defaults with this repo's own bucket functions), plus the host logic of `Newcode/NewLoadData.py` and the
`Train.evaluate_TopK / sample_negative` loops.  Every function cites the reference file:line it follows.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this module.  The product (`hhfm_b200/`) never does: it must fail loudly when the CUDA library is missing.

PARITY STATUS
  * loader, negative sampler, batch assembly, evaluate_TopK walk: PINNED against the reference's own code
    executed in the build container (tests/golden/make_reference_goldens.py imports /root/reference/Newcode
    with stub `tensorflow`/`toolz` modules and numpy/pandas compat shims; fixtures in tests/golden/).
  * model arithmetic (forward / loss / gradients / optimizers / top-k): **parity unpinned** -- the arithmetic
    lives in TensorFlow 1.x (unpinned, ~1.5-1.8, not vendored, not installable here: no TF-1 wheel for
    Python 3.12).  The restatement encodes TF's documented op semantics:
      l2_loss(t)=sum(t^2)/2; l2_regularizer(s)(w)=s*sum(w^2)/2; Adagrad acc0=0.1, acc+=g^2, w-=lr*g/sqrt(acc)
      (no epsilon); sparse (IndexedSlices) grads are summed per unique row and only touched rows move; a
      sparse + dense gradient on one variable aggregates to dense; TF1 Adam's sparse path equals the dense
      update with zero rows; reduce_max splits its gradient equally among ties; top_k sorts descending with
      the lower index first among equals; softmax subtracts the row max.
    Analytic gradients here are cross-checked against torch.autograd (tests/test_oracle.py).
    What the reference DOES hold for this arithmetic is statistical: its shipped datasets and the HR / NDCG / AUC lines
    of its run log (result.txt).  The CUDA path that this oracle checks is run on those datasets with the reference
    defaults and must land inside bands derived from that log (scripts/reference_bands.py,
    tests/test_gpu_reference_bands.py: HHFM frappe HR@5 0.67-0.69 against the reference's 0.674-0.698) -- a pin of the
    training outcome, not of single operations.

All arithmetic is float32 (np.float32 arrays; each NumPy ufunc rounds to fp32 like a non-fused GPU op).
Canonical score order for the top-N evaluators: products and sums are taken for k = 0..K-1 in ascending
order with a separately rounded multiply and add (no FMA) -- the CUDA exact rescoring kernel uses
`__fmul_rn/__fadd_rn` in the same order, so scores are bit-identical and index lists compare exactly.
"""
from __future__ import annotations

import math
from collections import defaultdict

import numpy as np

F32 = np.float32

POOL_SUM, POOL_MAX, POOL_MEAN = 0, 1, 2


# ----------------------------------------------------------------------------------------------------
# data: Newcode/NewLoadData.py:16-62
# ----------------------------------------------------------------------------------------------------
class LoadDataOracle:
    """Restatement of `LoadData.__init__` (NewLoadData.py:16-62) on an in-memory token table.

    tokens: object/str array [rows, 1+F]; column 0 is the label, 1 the user, 2 the item.
    The shuffle (NewLoadData.py:39) draws from the legacy global NumPy RNG exactly like the reference, so
    `np.random.seed(s)` before construction reproduces the reference split.
    """

    def __init__(self, tokens, ratio=0.9):
        tokens = np.asarray(tokens, dtype=object)
        self.n_user = len(set(tokens[:, 1].tolist()))        # NewLoadData.py:22
        self.n_item = len(set(tokens[:, 2].tolist()))        # NewLoadData.py:23
        ids = {}
        for tok in tokens[:, 1:].T.reshape(-1):              # column-major first-seen ids, :29-33
            if tok not in ids:
                ids[tok] = len(ids)
        self.features_M = len(ids)                           # :35
        data = np.empty(tokens.shape, dtype=np.int64)
        data[:, 0] = tokens[:, 0].astype(np.int64)
        for c in range(1, tokens.shape[1]):
            data[:, c] = [ids[t] for t in tokens[:, c]]      # :34
        self.Total_data = data
        np.random.shuffle(data)                              # :39 (in place, global RNG)
        test_size = int(len(data) * (1 - ratio))             # :40
        train, test = [], []
        self.positive_feedback = defaultdict(set)
        self.train_set = defaultdict(set)
        seen = set()
        i = 0
        cols = [c for c in range(1, data.shape[1]) if c != 2]
        for line in data:                                    # :48-58
            key = tuple(line[cols])
            if key not in seen and i < test_size:
                seen.add(key)
                test.append(line)
                i += 1
            else:
                self.positive_feedback[key].add(line[2])
                train.append(line)
                self.train_set[line[1]].add(line[2])
        self.Train_data = np.array(train)
        self.Test_data = np.array(test)


def sample_negative(rows, n_user, n_item, positive_feedback, num, randint=None):
    """`Train.sample_negative` (FM.py:284-294): uniform item draw with rejection against the positives of
    the row's key (all columns but the item).  `rows` = [n, F] ids without the label column."""
    randint = randint or np.random.randint
    samples = randint(n_user, n_user + n_item, size=(len(rows), num))
    cols = [c for c in range(rows.shape[1]) if c != 1]
    for i, row in enumerate(rows):
        key = tuple(row[cols])
        for j in range(num):
            neg = samples[i, j]
            while neg in positive_feedback[key]:
                samples[i, j] = neg = randint(n_user, n_user + n_item)
    return samples


def _splitmix64(x):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def sample_negative_hashed(key_id, num, n_user, n_item, pf_codes, span, seed):
    """Restatement of the DEVICE sampler (csrc/sampler.cu): the reference's uniform draw + rejection
    (FM.py:284-294) with a counter-based generator: draw(cell, attempt) = n_user + ((h >> 32) * n_item >> 32),
    h = splitmix64(seed ^ splitmix64(cell * 0x100000001B3 + attempt)), cell = row*num + column; a draw is rejected
    while key_id[row]*span + item is in pf_codes.  Bit-exact against the kernel (integer work)."""
    key_id = np.asarray(key_id, dtype=np.int64)
    n = len(key_id)
    codes = set(np.asarray(pf_codes, dtype=np.int64).tolist())
    out = np.zeros((n, num), dtype=np.int32)
    with np.errstate(over="ignore"):
        cell = np.arange(n * num, dtype=np.uint64)
        pending = np.ones(n * num, dtype=bool)
        attempt = 0
        flat = out.reshape(-1)
        kid = np.repeat(key_id, num)
        while pending.any() and attempt <= 4096:
            idx = np.flatnonzero(pending)
            h = _splitmix64(np.uint64(seed) ^ _splitmix64(cell[idx] * np.uint64(0x100000001B3) + np.uint64(attempt)))
            item = n_user + (((h >> np.uint64(32)) * np.uint64(n_item)) >> np.uint64(32)).astype(np.int64)
            flat[idx] = item
            if attempt >= 4096:
                break
            rej = np.array([(k >= 0) and ((k * span + it) in codes) for k, it in zip(kid[idx].tolist(), item.tolist())], dtype=bool)
            pending[idx[~rej]] = False
            attempt += 1
    return out


def assemble_pointwise_epoch(train, neg_samples, NG, neg_label):
    """FM.py:240-249 / AFM.py:309-318: positives followed by NG item-replaced copies labelled `neg_label`
    (FM: -0 == 0, AFM/DFM/MF: -1).  The shuffle (FM.py:250) is left to the caller."""
    copy_ = np.tile(train[:, None, :], [1, NG, 1]).reshape(-1, train.shape[1])
    copy_[:, 2] = neg_samples.reshape(-1)
    copy_[:, 0] = neg_label
    return np.append(train, copy_, axis=0)


# ----------------------------------------------------------------------------------------------------
# metrics: FM.py:325-359 (same text in AFM/DFM/BPR/OurModel7/CARS2)
# ----------------------------------------------------------------------------------------------------
def evaluate_topk_walk(prediction, rows, positive_feedback, TopK):
    """The per-row walk of `Train.evaluate_TopK` (FM.py:336-357), verbatim including its quirk: the
    `elif item in positive_feedback[key]` branch tests the *target* item, so such rows never advance `n`.

    prediction: int [C, tp] GLOBAL item ids (topk output + n_user, FM.py:335); rows: int [C, F].
    Returns the three result lists (hit, ndcg, reciprocal rank); rows that fall through append nothing.
    """
    res_map, res_ndcg, res_pre = [], [], []
    cols = [c for c in range(rows.shape[1]) if c != 1]
    for i, line in enumerate(rows):
        item = line[1]
        key = tuple(line[cols])
        n = 0
        for it in prediction[i]:
            if n > TopK - 1:
                res_map.append(0); res_ndcg.append(0); res_pre.append(0)
                break
            elif it == item:
                res_map.append(1)
                res_ndcg.append(np.log(2) / np.log(n + 2))
                res_pre.append(1 / (n + 1))
                break
            elif item in positive_feedback[key]:
                continue
            else:
                n = n + 1
    return res_map, res_ndcg, res_pre


def evaluate_topk_simple(prediction, items):
    """MF.py:265-274 / WDMF.py:247-256: hit iff the target is anywhere in the retrieved list."""
    return [1 if items[i] in prediction[i] else 0 for i in range(len(items))]


def topk_lowest_index(score, tp):
    """`tf.nn.top_k(score, tp)` indices: descending score, lower index first among equals."""
    score = np.asarray(score)
    order = np.lexsort((np.arange(score.shape[1])[None, :].repeat(score.shape[0], 0), -score), axis=1)
    return order[:, :tp].astype(np.int32)


# ----------------------------------------------------------------------------------------------------
# small fp32 helpers
# ----------------------------------------------------------------------------------------------------
def _f32(x):
    return np.asarray(x, dtype=F32)


def _sigmoid(x):
    x = _f32(x)
    return (F32(1.0) / (F32(1.0) + np.exp(-x))).astype(F32)


def seq_sum(a, axis):
    """Sequential fp32 sum along `axis` (index 0,1,2,... order) -- the canonical order of the CUDA kernels
    where the count is small (fields of one sample)."""
    a = np.moveaxis(_f32(a), axis, 0)
    acc = a[0].copy()
    for t in range(1, a.shape[0]):
        acc = (acc + a[t]).astype(F32)
    return acc


def seq_dot(a, b):
    """sum_k fl(a_k*b_k) for k ascending, fp32, multiply and add rounded separately (no FMA).
    a: [..., K], b: [..., K] broadcastable."""
    a = _f32(a); b = _f32(b)
    K = a.shape[-1]
    acc = (a[..., 0] * b[..., 0]).astype(F32)
    for k in range(1, K):
        acc = (acc + (a[..., k] * b[..., k]).astype(F32)).astype(F32)
    return acc


# ----------------------------------------------------------------------------------------------------
# FM: FM.py:99-126
# ----------------------------------------------------------------------------------------------------
def fm_forward(X, V, b, b0, val=None):
    """out[B] = sum_k 0.5((sum_f e_f)^2 - sum_f e_f^2) + sum_f b[x_f] + b0   (FM.py:99-120).
    e_f = val_f * V[x_f]; the reference's values are implicitly 1.0."""
    E = _f32(V)[X]                                       # :99  [B,F,K]
    if val is not None:
        E = (E * _f32(val)[:, :, None]).astype(F32)
    S = E.sum(axis=1, dtype=F32)                         # :100
    Q = (E * E).astype(F32).sum(axis=1, dtype=F32)       # :105-106
    fm = (F32(0.5) * (S * S - Q)).astype(F32)            # :109
    bil = fm.sum(axis=1, dtype=F32)                      # :113,117
    fb = _f32(b).reshape(-1)[X]
    if val is not None:
        fb = (fb * _f32(val)).astype(F32)
    fb = fb.sum(axis=1, dtype=F32)                       # :118
    out = (bil + fb + F32(b0)).astype(F32)               # :119-120
    return out, E, S


def fm_loss_grads(X, Y, V, b, b0, lamda=0.0, val=None):
    """Squared loss (FM.py:123-126) and its gradients (TF autodiff of the graph above).
    Returns loss, out, dV (dense [M,K]; includes +lamda*V when lamda>0), db [M], db0, and the list of
    touched rows.  Duplicate ids inside a row and across rows are summed (UnsortedSegmentSum)."""
    V = _f32(V); Y = _f32(Y).reshape(-1)
    out, E, S = fm_forward(X, V, b, b0, val)
    diff = (Y - out).astype(F32)
    loss = F32(0.5) * (diff * diff).astype(F32).sum(dtype=F32)
    g = (-diff).astype(F32)                              # d loss / d out
    if val is None:
        dE = (g[:, None, None] * (S[:, None, :] - E)).astype(F32)
        dbe = np.broadcast_to(g[:, None], X.shape).astype(F32)
    else:
        vv = _f32(val)[:, :, None]
        dE = (g[:, None, None] * vv * (S[:, None, :] - E)).astype(F32)
        dbe = (g[:, None] * _f32(val)).astype(F32)
    dV = np.zeros_like(V)
    np.add.at(dV, X.reshape(-1), dE.reshape(-1, V.shape[1]))
    db = np.zeros(V.shape[0], F32)
    np.add.at(db, X.reshape(-1), dbe.reshape(-1))
    db0 = g.sum(dtype=F32)
    if lamda > 0:
        loss = F32(loss + F32(lamda) * F32(0.5) * (V * V).astype(F32).sum(dtype=F32))   # :124
        dV = (dV + F32(lamda) * V).astype(F32)
    return F32(loss), out, dV, db, F32(db0), np.unique(X)


# ----------------------------------------------------------------------------------------------------
# MF: MF.py:81-98  (dropout keep=1 => identity; the bias term is computed but not added, :91-92)
# ----------------------------------------------------------------------------------------------------
def mf_forward(X, V):
    u = _f32(V)[X[:, 0]]; it = _f32(V)[X[:, 1]]
    return (u * it).astype(F32).sum(axis=1, dtype=F32), u, it


def mf_loss_grads(X, Y, V, lamda=0.0):
    V = _f32(V); Y = _f32(Y).reshape(-1)
    out, u, it = mf_forward(X, V)
    diff = (Y - out).astype(F32)
    loss = F32(0.5) * (diff * diff).astype(F32).sum(dtype=F32)
    g = (-diff).astype(F32)
    dV = np.zeros_like(V)
    np.add.at(dV, X[:, 0], (g[:, None] * it).astype(F32))
    np.add.at(dV, X[:, 1], (g[:, None] * u).astype(F32))
    if lamda > 0:
        loss = F32(loss + F32(lamda) * F32(0.5) * (V * V).astype(F32).sum(dtype=F32))
        dV = (dV + F32(lamda) * V).astype(F32)
    return F32(loss), out, dV


# ----------------------------------------------------------------------------------------------------
# dropout on the interaction layer (FM.py:114, MF.py:87): tf.nn.dropout(x, keep) = x / keep * floor(keep + u)
# ----------------------------------------------------------------------------------------------------
def dropout_mask_hashed(seed, B, K, keep):
    """Restatement of the DEVICE dropout mask (csrc/common.cuh dropout_keep01): element (s, k) is kept iff
    (splitmix64(seed ^ splitmix64(s * 0x100000001B3 + k)) >> 40) * 2^-24 < keep.  Returns the 0/1 mask float32 [B, K].
    TF draws its mask from its own stream; this is statistically the same dropout, bit-exact against the kernel."""
    with np.errstate(over="ignore"):
        s = np.arange(B, dtype=np.uint64)[:, None]
        k = np.arange(K, dtype=np.uint64)[None, :]
        h = _splitmix64(np.uint64(seed) ^ _splitmix64(s * np.uint64(0x100000001B3) + k))
    u = (h >> np.uint64(40)).astype(np.float32) * F32(1.0 / 16777216.0)
    return (u < F32(keep)).astype(F32)


def fm_dropout_loss_grads(X, Y, V, b, b0, keep, seed, lamda=0.0):
    """FM.py:99-126 with dropout_keep < 1 (:114): FM = 0.5*(S^2 - sum e^2); FM = FM / keep * mask; out = sum_k FM + ..."""
    V = _f32(V); Y = _f32(Y).reshape(-1)
    E = V[X]
    S = E.sum(axis=1, dtype=F32)
    fm = (F32(0.5) * ((S * S).astype(F32) - (E * E).astype(F32).sum(axis=1, dtype=F32))).astype(F32)
    mk = (dropout_mask_hashed(seed, X.shape[0], V.shape[1], keep) / F32(keep)).astype(F32)
    fm = ((fm / F32(keep)).astype(F32) * dropout_mask_hashed(seed, X.shape[0], V.shape[1], keep)).astype(F32)
    out = fm.sum(axis=1, dtype=F32)
    if b is not None:
        out = (out + _f32(b).reshape(-1)[X].sum(axis=1, dtype=F32)).astype(F32)
    out = (out + F32(b0 if b0 is not None else 0.0)).astype(F32)
    diff = (Y - out).astype(F32)
    loss = F32(0.5) * (diff * diff).astype(F32).sum(dtype=F32)
    g = (-diff).astype(F32)
    dE = ((g[:, None, None] * (S[:, None, :] - E)).astype(F32) * mk[:, None, :]).astype(F32)
    dV = np.zeros_like(V)
    np.add.at(dV, X.reshape(-1), dE.reshape(-1, V.shape[1]))
    db = np.zeros(V.shape[0], F32)
    np.add.at(db, X.reshape(-1), np.broadcast_to(g[:, None], X.shape).reshape(-1))
    if lamda > 0:
        loss = F32(loss + F32(lamda) * F32(0.5) * (V * V).astype(F32).sum(dtype=F32))
        dV = (dV + F32(lamda) * V).astype(F32)
    return F32(loss), out, dV, db, F32(g.sum(dtype=F32))


def mf_dropout_loss_grads(X, Y, V, keep, seed, lamda=0.0):
    """MF.py:81-98 with the reference default dropout_keep = 0.7 (:31,87)."""
    V = _f32(V); Y = _f32(Y).reshape(-1)
    u = V[X[:, 0]]; it = V[X[:, 1]]
    m01 = dropout_mask_hashed(seed, X.shape[0], V.shape[1], keep)
    mk = (m01 / F32(keep)).astype(F32)
    out = (((u * it).astype(F32) / F32(keep)).astype(F32) * m01).astype(F32).sum(axis=1, dtype=F32)
    diff = (Y - out).astype(F32)
    loss = F32(0.5) * (diff * diff).astype(F32).sum(dtype=F32)
    g = (-diff).astype(F32)
    dV = np.zeros_like(V)
    np.add.at(dV, X[:, 0], ((g[:, None] * it).astype(F32) * mk).astype(F32))
    np.add.at(dV, X[:, 1], ((g[:, None] * u).astype(F32) * mk).astype(F32))
    if lamda > 0:
        loss = F32(loss + F32(lamda) * F32(0.5) * (V * V).astype(F32).sum(dtype=F32))
        dV = (dV + F32(lamda) * V).astype(F32)
    return F32(loss), out, dV


# ----------------------------------------------------------------------------------------------------
# HHFM (OurModel7.py:105-184) and BPR (BPR.py:76-88): pairwise ranking with max-negative
# ----------------------------------------------------------------------------------------------------
def _pool(E, mode):
    """Pooling1*/(tf.reduce_sum|reduce_max|reduce_mean)(E, axis=1) (OurModel7.py:14-19)."""
    if mode == POOL_SUM:
        return E.sum(axis=1, dtype=F32)
    if mode == POOL_MEAN:
        return E.mean(axis=1, dtype=F32)
    return E.max(axis=1)


def _pool_bwd(E, out, d_out, mode):
    """Gradient of _pool w.r.t. E [B,n,K].  reduce_max splits equally among ties (TF _MinOrMaxGrad)."""
    n = E.shape[1]
    if mode == POOL_SUM:
        return np.broadcast_to(d_out[:, None, :], E.shape).astype(F32)
    if mode == POOL_MEAN:
        return np.broadcast_to((d_out / F32(n))[:, None, :], E.shape).astype(F32)
    ind = (E == out[:, None, :]).astype(F32)
    cnt = ind.sum(axis=1, keepdims=True, dtype=F32)
    return (ind / cnt * d_out[:, None, :]).astype(F32)


def hybrid_feature(V, Pos, Fea=None, Tim=None, pools=(POOL_SUM, POOL_SUM, POOL_SUM)):
    """hyb = Pool1F(stack[users, Pool1C(V[Fea]), Pool1T(V[Tim])]) (OurModel7.py:105-168).  The pooled pair
    products (Pool2*) are built but commented out of the sums in the reference (:124,141,168) -> omitted.
    BPR (BPR.py:76) is the special case Fea=Tim=None: hyb = V[user]."""
    V = _f32(V)
    parts = {"u": V[Pos[:, 0]]}
    stack = [parts["u"]]
    if Fea is not None and Fea.shape[1] > 0:
        parts["EC"] = V[Fea]; parts["C"] = _pool(parts["EC"], pools[0]); stack.append(parts["C"])
    if Tim is not None and Tim.shape[1] > 0:
        parts["ET"] = V[Tim]; parts["T"] = _pool(parts["ET"], pools[1]); stack.append(parts["T"])
    parts["stack"] = np.stack(stack, axis=1)             # [B,num,K]
    if len(stack) == 1:
        hyb = stack[0]
        parts["single"] = True
    else:
        hyb = _pool(parts["stack"], pools[2])
    return hyb.astype(F32), parts


def pairrank_scores(V, Pos, Neg, Fea=None, Tim=None, pools=(0, 0, 0)):
    """PositiveFeadback [B] and NegativeFeadback [B,NG] (OurModel7.py:171-172, BPR.py:79-80)."""
    V = _f32(V)
    hyb, parts = hybrid_feature(V, Pos, Fea, Tim, pools)
    vp = V[Pos[:, 1]]
    pos = (hyb * vp).astype(F32).sum(axis=1, dtype=F32)
    if Neg is None:
        return pos, None, hyb, parts
    vn = V[Neg]
    neg = (hyb[:, None, :] * vn).astype(F32).sum(axis=2, dtype=F32)
    return pos, neg, hyb, parts


def pairrank_loss_grads(V, Pos, Neg, Fea=None, Tim=None, pools=(0, 0, 0), lamda=0.0):
    """loss = -sum log(sigmoid(pos - max_j neg_j)) (+ lamda*0.5*||V||^2)  (OurModel7.py:174-182, BPR.py:81-86)
    and the dense gradient dV.  Also returns pos, neg."""
    V = _f32(V)
    pos, neg, hyb, parts = pairrank_scores(V, Pos, Neg, Fea, Tim, pools)
    m = neg.max(axis=1)
    x = (pos - m).astype(F32)
    sg = _sigmoid(x)
    loss = -np.log(sg).astype(F32).sum(dtype=F32)
    g = (sg - F32(1.0)).astype(F32)                      # d loss / d x  = -(1 - sigmoid(x))
    tie = (neg == m[:, None]).astype(F32)
    dneg = (-g[:, None] * tie / tie.sum(axis=1, keepdims=True, dtype=F32)).astype(F32)
    vp = V[Pos[:, 1]]; vn = V[Neg]
    dhyb = (g[:, None] * vp + (dneg[:, :, None] * vn).astype(F32).sum(axis=1, dtype=F32)).astype(F32)
    dV = np.zeros_like(V)
    np.add.at(dV, Pos[:, 1], (g[:, None] * hyb).astype(F32))
    np.add.at(dV, Neg.reshape(-1), (dneg[:, :, None] * hyb[:, None, :]).astype(F32).reshape(-1, V.shape[1]))
    if parts.get("single"):
        np.add.at(dV, Pos[:, 0], dhyb)
    else:
        dstack = _pool_bwd(parts["stack"], hyb, dhyb, pools[2])
        s = 0
        np.add.at(dV, Pos[:, 0], dstack[:, s]); s += 1
        if "C" in parts:
            dEC = _pool_bwd(parts["EC"], parts["C"], dstack[:, s], pools[0]); s += 1
            np.add.at(dV, Fea.reshape(-1), dEC.reshape(-1, V.shape[1]))
        if "T" in parts:
            dET = _pool_bwd(parts["ET"], parts["T"], dstack[:, s], pools[1]); s += 1
            np.add.at(dV, Tim.reshape(-1), dET.reshape(-1, V.shape[1]))
    if lamda > 0:
        loss = F32(loss + F32(lamda) * F32(0.5) * (V * V).astype(F32).sum(dtype=F32))
        dV = (dV + F32(lamda) * V).astype(F32)
    return F32(loss), pos, neg, dV


# ----------------------------------------------------------------------------------------------------
# AFM: AFM.py:103-148
# ----------------------------------------------------------------------------------------------------
def pair_index(F):
    """(i,j), i<j in the lexicographic order of AFM.py:107-110."""
    return [(i, j) for i in range(F) for j in range(i + 1, F)]


def afm_forward(X, w):
    """w: dict feature_embeddings [M,K], feature_bias [M,1], bias, attention_W [K,A], attention_b [1,A],
    attention_p [A], prediction [K,1].  Returns out [B] and a cache for the backward."""
    V = _f32(w["feature_embeddings"])
    E = V[X]                                                             # :103
    pi = pair_index(X.shape[1])
    I = np.array([p[0] for p in pi]); J = np.array([p[1] for p in pi])
    P = (E[:, I, :] * E[:, J, :]).astype(F32)                            # :105-112  [B,P,K]
    W = _f32(w["attention_W"]); ab = _f32(w["attention_b"]).reshape(-1); ap = _f32(w["attention_p"]).reshape(-1)
    Z = (P.reshape(-1, P.shape[2]) @ W).astype(F32).reshape(P.shape[0], P.shape[1], -1) + ab   # :117-118,123
    H = np.maximum(Z, F32(0)).astype(F32)
    s = (H * ap).astype(F32).sum(axis=2, dtype=F32)                      # :123-124  [B,P]
    smax = s.max(axis=1, keepdims=True)
    ex = np.exp((s - smax).astype(F32)).astype(F32)
    a = (ex / ex.sum(axis=1, keepdims=True, dtype=F32)).astype(F32)      # :125 softmax over pairs
    afm = (a[:, :, None] * P).astype(F32).sum(axis=1, dtype=F32)         # :130  [B,K]
    wp = _f32(w["prediction"]).reshape(-1)
    bil = (afm * wp).astype(F32).sum(axis=1, dtype=F32)                  # :138-139
    fb = _f32(w["feature_bias"]).reshape(-1)[X].sum(axis=1, dtype=F32)   # :140
    out = (bil + fb + F32(np.asarray(w["bias"]).reshape(()))).astype(F32)  # :141-142
    return out, dict(E=E, P=P, Z=Z, H=H, a=a, afm=afm, I=I, J=J)


def afm_loss_grads(X, Y, w, lamda_attention=0.0):
    """0.5*sum(y-out)^2 + lamda_attention*0.5*||attention_W||^2 (AFM.py:146) and all gradients."""
    Y = _f32(Y).reshape(-1)
    out, c = afm_forward(X, w)
    V = _f32(w["feature_embeddings"]); W = _f32(w["attention_W"])
    ap = _f32(w["attention_p"]).reshape(-1); wp = _f32(w["prediction"]).reshape(-1)
    diff = (Y - out).astype(F32)
    loss = F32(0.5) * (diff * diff).astype(F32).sum(dtype=F32)
    g = (-diff).astype(F32)
    B, Pn, K = c["P"].shape
    d_afm = (g[:, None] * wp[None, :]).astype(F32)                       # [B,K]
    d_wp = (g[:, None] * c["afm"]).astype(F32).sum(axis=0, dtype=F32)
    d_a = (c["P"] * d_afm[:, None, :]).astype(F32).sum(axis=2, dtype=F32)      # [B,P]
    dP = (c["a"][:, :, None] * d_afm[:, None, :]).astype(F32)
    d_s = (c["a"] * (d_a - (c["a"] * d_a).astype(F32).sum(axis=1, keepdims=True, dtype=F32))).astype(F32)
    d_ap = (d_s[:, :, None] * c["H"]).astype(F32).sum(axis=(0, 1), dtype=F32)
    d_Z = (d_s[:, :, None] * ap[None, None, :] * (c["Z"] > 0)).astype(F32)       # [B,P,A]
    d_ab = d_Z.sum(axis=(0, 1), dtype=F32)
    d_W = (c["P"].reshape(-1, K).T @ d_Z.reshape(B * Pn, -1)).astype(F32)
    dP = (dP + (d_Z.reshape(B * Pn, -1) @ W.T).astype(F32).reshape(B, Pn, K)).astype(F32)
    dE = np.zeros_like(c["E"])
    np.add.at(dE, (slice(None), c["I"]), (dP * c["E"][:, c["J"], :]).astype(F32))
    np.add.at(dE, (slice(None), c["J"]), (dP * c["E"][:, c["I"], :]).astype(F32))
    dV = np.zeros_like(V)
    np.add.at(dV, X.reshape(-1), dE.reshape(-1, K))
    db = np.zeros(V.shape[0], F32)
    np.add.at(db, X.reshape(-1), np.broadcast_to(g[:, None], X.shape).reshape(-1))
    if lamda_attention > 0:
        loss = F32(loss + F32(lamda_attention) * F32(0.5) * (W * W).astype(F32).sum(dtype=F32))
        d_W = (d_W + F32(lamda_attention) * W).astype(F32)
    grads = dict(feature_embeddings=dV, feature_bias=db.reshape(-1, 1), bias=g.sum(dtype=F32),
                 attention_W=d_W, attention_b=d_ab.reshape(1, -1), attention_p=d_ap, prediction=d_wp.reshape(-1, 1))
    return F32(loss), out, grads


# ----------------------------------------------------------------------------------------------------
# DeepFM: DFM.py:104-152
# ----------------------------------------------------------------------------------------------------
def dfm_forward(X, w, n_layers=3):
    V = _f32(w["feature_embeddings"])
    E = V[X]                                                             # :104-105
    y1 = _f32(w["feature_bias"]).reshape(-1)[X]                          # :109-110  [B,F]
    S = E.sum(axis=1, dtype=F32)
    y2 = (F32(0.5) * (S * S - (E * E).astype(F32).sum(axis=1, dtype=F32))).astype(F32)   # :114-122
    h = E.reshape(E.shape[0], -1)                                        # :125
    acts = [h]
    for i in range(n_layers):                                            # :126-128
        h = np.maximum((h @ _f32(w["layer_%d" % i])).astype(F32) + _f32(w["bias_%d" % i]).reshape(1, -1), F32(0)).astype(F32)
        acts.append(h)
    cat = np.concatenate([y1, y2, h], axis=1)                            # :131-132
    out = ((cat @ _f32(w["concat_projection"])).astype(F32).reshape(-1) + F32(np.asarray(w["concat_bias"]).reshape(()))).astype(F32)  # :137
    return out, dict(E=E, S=S, acts=acts, cat=cat)


def dfm_loss_grads(X, Y, w, l2_reg=0.0, n_layers=3):
    """DFM.py:139-152: loss = 0.5*sum(y-out)^2 + l2_reg/2*(||concat_projection||^2 + sum_i ||layer_i||^2) and the
    gradient of every variable (TF autodiff restated; relu'(0) = 0)."""
    Y = _f32(Y).reshape(-1)
    out, c = dfm_forward(X, w, n_layers)
    V = _f32(w["feature_embeddings"])
    B, F, K = c["E"].shape
    proj = _f32(w["concat_projection"]).reshape(-1)
    diff = (Y - out).astype(F32)
    loss = F32(0.5) * (diff * diff).astype(F32).sum(dtype=F32)
    g = (-diff).astype(F32)                                               # d loss / d out
    grads = {}
    grads["concat_projection"] = (c["cat"].T @ g).astype(F32).reshape(-1, 1)
    grads["concat_bias"] = g.sum(dtype=F32)
    d_cat = (g[:, None] * proj[None, :]).astype(F32)
    d_y1, d_y2, d_h = d_cat[:, :F], d_cat[:, F:F + K], d_cat[:, F + K:]
    for i in reversed(range(n_layers)):
        h_out, h_in = c["acts"][i + 1], c["acts"][i]
        dZ = (d_h * (h_out > 0)).astype(F32)
        grads["layer_%d" % i] = (h_in.T @ dZ).astype(F32)
        grads["bias_%d" % i] = dZ.sum(axis=0, dtype=F32).reshape(1, -1)
        d_h = (dZ @ _f32(w["layer_%d" % i]).T).astype(F32)
    dE = (d_h.reshape(B, F, K) + (d_y2[:, None, :] * (c["S"][:, None, :] - c["E"])).astype(F32)).astype(F32)
    dV = np.zeros_like(V)
    np.add.at(dV, X.reshape(-1), dE.reshape(-1, K))
    db = np.zeros(V.shape[0], F32)
    np.add.at(db, X.reshape(-1), d_y1.reshape(-1))
    grads["feature_embeddings"] = dV
    grads["feature_bias"] = db.reshape(-1, 1)
    if l2_reg > 0:
        for k in ["concat_projection"] + ["layer_%d" % i for i in range(n_layers)]:
            wk = _f32(w[k])
            loss = F32(loss + F32(l2_reg) * F32(0.5) * (wk * wk).astype(F32).sum(dtype=F32))
            grads[k] = (grads[k] + F32(l2_reg) * wk.reshape(grads[k].shape)).astype(F32)
    return F32(loss), out, grads


# ----------------------------------------------------------------------------------------------------
# Wide&Deep: WDMF.py:51-126 (tf.contrib.learn.DNNLinearCombinedClassifier).  The arithmetic of the reference lives inside
# TensorFlow (hashed / crossed / embedding columns, FTRL for the linear half, Adagrad for the DNN); it is restated here from
# TF's documented behaviour with our own bucket functions -- PARITY UNPINNED: nothing in the reference tree pins
# TF's string fingerprints, its column order or its initialisers.
# ----------------------------------------------------------------------------------------------------
def _splitmix64_arr(x):
    x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def wd_cross_bucket(xi, xj, n_buckets):
    """Bucket of the crossed column (x_i, x_j): splitmix64((x_i << 32) | x_j) mod n_buckets (crossed_column(..., 1e4),
    WDMF.py:66; TF's own fingerprint is not restatable)."""
    with np.errstate(over="ignore"):
        key = (np.asarray(xi, dtype=np.uint64) << np.uint64(32)) | np.asarray(xj, dtype=np.uint64)
        return (_splitmix64_arr(key) % np.uint64(n_buckets)).astype(np.int64)


def wd_wide(X, w):
    """b_wide + sum_f w_lin[x_f] + sum_{i<j} w_cross[p(i,j)][bucket(x_i, x_j)]  (linear_feature_columns, WDMF.py:57-66;
    single columns: the loader's global id is its own bucket).  Returns the logit [B] and the cross buckets [B,P]."""
    X = np.asarray(X)
    wl = _f32(w["wide_linear"]).reshape(-1); wc = _f32(w["wide_cross"])
    pi = pair_index(X.shape[1])
    buckets = np.stack([wd_cross_bucket(X[:, i], X[:, j], wc.shape[1]) for i, j in pi], axis=1) if pi else np.zeros((len(X), 0), np.int64)
    acc = np.zeros(len(X), F32)
    for f in range(X.shape[1]):
        acc = (acc + wl[X[:, f]]).astype(F32)
    for p in range(len(pi)):
        acc = (acc + wc[p, buckets[:, p]]).astype(F32)
    return (acc + F32(np.asarray(w["wide_bias"]).reshape(()))).astype(F32), buckets


def wd_forward(X, w, n_layers=3):
    """logit = wide + deep; deep = relu MLP over the concatenated 128-d embeddings (dnn_feature_columns, hidden units
    [1024, 512, 256], WDMF.py:67-73) and a [D_L,1] logits layer with bias."""
    wide, buckets = wd_wide(X, w)
    E = _f32(w["feature_embeddings"])[np.asarray(X)]
    h = E.reshape(E.shape[0], -1)
    acts = [h]
    for i in range(n_layers):
        h = np.maximum((h @ _f32(w["layer_%d" % i])).astype(F32) + _f32(w["bias_%d" % i]).reshape(1, -1), F32(0)).astype(F32)
        acts.append(h)
    deep = ((h @ _f32(w["logits_w"]).reshape(-1, 1)).astype(F32).reshape(-1) + F32(np.asarray(w["logits_b"]).reshape(()))).astype(F32)
    return (deep + wide).astype(F32), dict(E=E, acts=acts, buckets=buckets)


def wd_loss_grads(X, Y, w, n_layers=3):
    """Mean sigmoid cross-entropy over the batch (labels in {0,1}; the estimator's head) and the gradient of every variable."""
    X = np.asarray(X); Y = _f32(Y).reshape(-1)
    z, c = wd_forward(X, w, n_layers)
    B, F, K = c["E"].shape
    inv = F32(1.0 / B)
    per = (np.maximum(z, F32(0)) - z * Y + np.log1p(np.exp(-np.abs(z)))).astype(F32)
    loss = F32((per * inv).astype(F32).sum(dtype=F32))
    g = ((_sigmoid(z) - Y) * inv).astype(F32)                              # d loss / d logit
    grads = {}
    gl = np.zeros(_f32(w["wide_linear"]).reshape(-1).shape, F32)
    np.add.at(gl, X.reshape(-1), np.repeat(g, F))
    gc = np.zeros(_f32(w["wide_cross"]).shape, F32)
    for p in range(c["buckets"].shape[1]):
        np.add.at(gc[p], c["buckets"][:, p], g)
    grads["wide_linear"] = gl; grads["wide_cross"] = gc; grads["wide_bias"] = g.sum(dtype=F32)
    h_last = c["acts"][-1]
    grads["logits_w"] = (h_last.T @ g).astype(F32).reshape(-1, 1)
    grads["logits_b"] = g.sum(dtype=F32)
    d_h = (g[:, None] * _f32(w["logits_w"]).reshape(1, -1)).astype(F32)
    for i in reversed(range(n_layers)):
        h_out, h_in = c["acts"][i + 1], c["acts"][i]
        dZ = (d_h * (h_out > 0)).astype(F32)
        grads["layer_%d" % i] = (h_in.T @ dZ).astype(F32)
        grads["bias_%d" % i] = dZ.sum(axis=0, dtype=F32).reshape(1, -1)
        d_h = (dZ @ _f32(w["layer_%d" % i]).T).astype(F32)
    dV = np.zeros_like(_f32(w["feature_embeddings"]))
    np.add.at(dV, X.reshape(-1), d_h.reshape(-1, K))
    grads["feature_embeddings"] = dV
    return loss, z, grads


def ftrl_dense(w, accum, linear, g, lr, l1=0.0, l2=0.0):
    """TF1 ApplyFtrl, learning_rate_power = -0.5 (tf.train.FtrlOptimizer defaults: the linear half of
    DNNLinearCombinedClassifier): n' = n + g^2; z += g - (sqrt(n') - sqrt(n))/lr * w;
    w = (sign(z) l1 - z) / (sqrt(n')/lr + 2 l2) if |z| > l1 else 0."""
    w = _f32(w); accum = _f32(accum); linear = _f32(linear); g = _f32(g)
    a1 = (accum + g * g).astype(F32)
    z = (linear + g - ((np.sqrt(a1) - np.sqrt(accum)) / F32(lr) * w).astype(F32)).astype(F32)
    quad = (np.sqrt(a1) / F32(lr) + F32(2.0 * l2)).astype(F32)
    w1 = np.where(np.abs(z) > F32(l1), ((np.sign(z) * F32(l1) - z) / quad).astype(F32), F32(0)).astype(F32)
    return w1, a1, z


# ----------------------------------------------------------------------------------------------------
# CARS2: CARS2.py:85-123 (restated op by op; the device kernels use the collapsed form T = sum_q B_q Z[:,q,:])
# ----------------------------------------------------------------------------------------------------
def cars2_feedback(Pos, Fea, w, items=None):
    """PositiveFeadback (CARS2.py:104-106) for rows Pos [B,2] (user, item) and context ids Fea [B]."""
    UI = _f32(w["UI"]); C = _f32(w["Context"])[np.asarray(Fea)]
    u = UI[np.asarray(Pos)[:, 0]]
    it = UI[np.asarray(Pos)[:, 1]] if items is None else items
    pik = np.einsum("bd,dpc,bc->bp", u, _f32(w["W"]), C).astype(F32)                  # :92-93
    qjk = np.einsum("bd,dqc,bc->bq", it, _f32(w["Z"]), C).astype(F32)                 # :95-96
    return ((u * it).sum(axis=1, dtype=F32) + (pik * _f32(w["A"])).sum(axis=1, dtype=F32)
            + (qjk * _f32(w["B"])).sum(axis=1, dtype=F32)).astype(F32)


def cars2_loss_grads(Pos, Fea, Neg, w, lamda=0.0):
    """loss = -sum log sigmoid(Pos - Neg) + lamda/2 * sum ||.||^2 over the six variables (CARS2.py:113-123) and all
    gradients; Negitems is the SUM of the negative item rows (:90)."""
    Pos = np.asarray(Pos); Neg = np.asarray(Neg); Fea = np.asarray(Fea)
    UI = _f32(w["UI"]); Ctx = _f32(w["Context"]); W = _f32(w["W"]); Z = _f32(w["Z"]); A = _f32(w["A"]); Bv = _f32(w["B"])
    u = UI[Pos[:, 0]]; ip = UI[Pos[:, 1]]; ineg = UI[Neg].sum(axis=1, dtype=F32); C = Ctx[Fea]
    pos = cars2_feedback(Pos, Fea, w)
    neg = cars2_feedback(Pos, Fea, w, items=ineg)
    x = (pos - neg).astype(F32)
    sg = _sigmoid(x)
    loss = F32(-np.log(sg).astype(F32).sum(dtype=F32))
    g = (sg - F32(1.0)).astype(F32)                                        # d loss / d x
    d_pos, d_neg = g, -g
    # d u: from u.ip, u.ineg and pik (pik.A appears in both scores: its contributions cancel)
    pikA_u = np.einsum("dpc,p,bc->bd", W, A, C).astype(F32)
    du = (d_pos[:, None] * (ip + pikA_u) + d_neg[:, None] * (ineg + pikA_u)).astype(F32)
    ZB_c = np.einsum("dqc,q,bc->bd", Z, Bv, C).astype(F32)
    dip = (d_pos[:, None] * (u + ZB_c)).astype(F32)
    dineg = (d_neg[:, None] * (u + ZB_c)).astype(F32)
    dC = (np.einsum("b,bd,dpc,p->bc", d_pos + d_neg, u, W, A) + np.einsum("b,bd,dqc,q->bc", d_pos, ip, Z, Bv)
          + np.einsum("b,bd,dqc,q->bc", d_neg, ineg, Z, Bv)).astype(F32)
    dW = np.einsum("b,bd,p,bc->dpc", d_pos + d_neg, u, A, C).astype(F32)
    dA = np.einsum("b,bd,dpc,bc->p", d_pos + d_neg, u, W, C).astype(F32)
    dZ = (np.einsum("b,bd,q,bc->dqc", d_pos, ip, Bv, C) + np.einsum("b,bd,q,bc->dqc", d_neg, ineg, Bv, C)).astype(F32)
    dB = (np.einsum("b,bd,dqc,bc->q", d_pos, ip, Z, C) + np.einsum("b,bd,dqc,bc->q", d_neg, ineg, Z, C)).astype(F32)
    dUI = np.zeros_like(UI); dCtx = np.zeros_like(Ctx)
    np.add.at(dUI, Pos[:, 0], du); np.add.at(dUI, Pos[:, 1], dip)
    for j in range(Neg.shape[1]):
        np.add.at(dUI, Neg[:, j], dineg)
    np.add.at(dCtx, Fea, dC)
    grads = dict(UI=dUI, Context=dCtx, W=dW, Z=dZ, A=dA, B=dB)
    if lamda > 0:
        for k in grads:
            wk = _f32(w[k])
            loss = F32(loss + F32(lamda) * F32(0.5) * (wk * wk).astype(F32).sum(dtype=F32))
            grads[k] = (grads[k] + F32(lamda) * wk).astype(F32)
    return F32(loss), pos, grads


def cars2_topk_scores(users, Fea, w, n_user, n_item):
    """Feedback [C, N] of CARS2.topk (CARS2.py:171-187)."""
    UI = _f32(w["UI"]); C = _f32(w["Context"])[np.asarray(Fea)]
    u = UI[np.asarray(users)]; items = UI[n_user:n_user + n_item]
    pik = np.einsum("bd,dpc,bc->bp", u, _f32(w["W"]), C).astype(F32)
    qjk = np.einsum("nd,dqc,bc->bnq", items, _f32(w["Z"]), C).astype(F32)
    return ((u @ items.T).astype(F32) + (pik * _f32(w["A"])).sum(axis=1, keepdims=True, dtype=F32)
            + (qjk * _f32(w["B"])).sum(axis=2, dtype=F32)).astype(F32)


# ----------------------------------------------------------------------------------------------------
# optimizers: TF1 semantics (FM.py:129-136, BPR.py:93, MF.py:104)
# ----------------------------------------------------------------------------------------------------
def adagrad_dense(w, acc, g, lr):
    """ApplyAdagrad: acc += g^2; w -= lr * g / sqrt(acc)   (no epsilon; acc0 = 0.1 or 1e-8)."""
    acc = (_f32(acc) + (_f32(g) * _f32(g)).astype(F32)).astype(F32)
    w = (_f32(w) - (F32(lr) * _f32(g) / np.sqrt(acc)).astype(F32)).astype(F32)
    return w, acc


def adagrad_rows(w, acc, g, rows, lr):
    """SparseApplyAdagrad on the de-duplicated rows: untouched rows keep w and acc."""
    w = _f32(w).copy(); acc = _f32(acc).copy()
    w[rows], acc[rows] = adagrad_dense(w[rows], acc[rows], _f32(g)[rows], lr)
    return w, acc


def adagrad_dense_l2(w, acc, g, lr, lamda):
    """The reference's dense step on a regularised table (FM.py:124,132): g_eff = g + lamda * w, then ApplyAdagrad."""
    ge = (_f32(g) + (F32(lamda) * _f32(w)).astype(F32)).astype(F32)
    return adagrad_dense(w, acc, ge, lr)


def adagrad_l2_lazy_replay(w, acc, last_step, rows, upto, lr, lamda):
    """Lazy-exact dense L2 (hard part 2-ii; csrc/opt.cu adagrad_l2_replay_kernel): a row that no batch gathered at
    steps last_step+1 .. upto saw g = 0 there, i.e. g_eff = lamda * w, a recurrence of the row alone.  Replays those steps for
    `rows` (None = every row) and stamps them with `upto`; the result has the bits the dense schedule would have produced."""
    w = _f32(w).copy(); acc = _f32(acc).copy(); last_step = np.asarray(last_step).copy()
    rows = np.arange(len(w)) if rows is None else np.asarray(rows)
    for r in rows:
        for _ in range(int(upto) - int(last_step[r])):
            ge = (F32(lamda) * w[r]).astype(F32)
            acc[r] = (acc[r] + (ge * ge).astype(F32)).astype(F32)
            w[r] = (w[r] - (F32(lr) * ge / np.sqrt(acc[r])).astype(F32)).astype(F32)
        last_step[r] = upto
    return w, acc, last_step


def adam_dense(w, m, v, g, lr, t, beta1=0.9, beta2=0.999, eps=1e-8):
    """TF1 AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); var -= lr_t*m/(sqrt(v)+eps).  The sparse path
    (`_apply_sparse_shared`) decays m,v of every row and moves every row, i.e. equals this with zero rows."""
    lr_t = F32(lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t))
    g = _f32(g)
    m = (F32(beta1) * _f32(m) + F32(1 - beta1) * g).astype(F32)
    v = (F32(beta2) * _f32(v) + F32(1 - beta2) * (g * g).astype(F32)).astype(F32)
    w = (_f32(w) - (lr_t * m / (np.sqrt(v) + F32(eps))).astype(F32)).astype(F32)
    return w, m, v


def momentum_dense(w, acc, g, lr, momentum=0.95):
    """ApplyMomentum: acc = acc*momentum + g; w -= lr*acc."""
    acc = (_f32(acc) * F32(momentum) + _f32(g)).astype(F32)
    return (_f32(w) - F32(lr) * acc).astype(F32), acc


def momentum_rows(w, acc, g, rows, lr, momentum=0.95):
    w = _f32(w).copy(); acc = _f32(acc).copy()
    w[rows], acc[rows] = momentum_dense(w[rows], acc[rows], _f32(g)[rows], lr, momentum)
    return w, acc


def sgd_dense(w, g, lr):
    return (_f32(w) - F32(lr) * _f32(g)).astype(F32)


# ----------------------------------------------------------------------------------------------------
# full-catalog top-N: canonical fp32 scores + lowest-index top_k
# ----------------------------------------------------------------------------------------------------
def fm_topk_scores(A, V, b, n_user, n_item):
    """FM.topk (FM.py:174-185): score[c,n] = b[n] + sum_k (u+Fc)_k * (v_n+Fc)_k, Fc = sum of ctx rows.
    Canonical order: Fc sequential over columns 2..; A_k = fl(u_k+Fc_k); B_nk = fl(v_nk+Fc_k); seq_dot;
    then fl(bias + score)."""
    V = _f32(V)
    u = V[A[:, 0]]
    if A.shape[1] > 2:
        Fc = seq_sum(V[A[:, 2:]], axis=1)
    else:
        Fc = np.zeros_like(u)
    UF = (u + Fc).astype(F32)                                   # :177
    items = V[n_user:n_user + n_item]
    IF = (items[None, :, :] + Fc[:, None, :]).astype(F32)       # :178
    score = seq_dot(UF[:, None, :], IF)                         # :180-183
    bias = _f32(b).reshape(-1)[n_user:n_user + n_item]
    return (bias[None, :] + score).astype(F32)                  # :185


def dot_topk_scores(Q, V, n_user, n_item):
    """BPR.topk / MF.topk / OUR.topk (BPR.py:132-135, MF.py:145-148, OurModel7.py:294): score = q_c . v_n."""
    items = _f32(V)[n_user:n_user + n_item]
    Q = _f32(Q)
    if Q.shape[0] * items.shape[0] < (1 << 22):
        return seq_dot(Q[:, None, :], items[None, :, :])
    # the same arithmetic as seq_dot (k ascending, multiply and add rounded separately) streamed over a K-major copy of the
    # catalog: 10x faster for the million-item checks of tests/test_gpu_fullsize.py
    items_t = np.ascontiguousarray(items.T)
    acc = (Q[:, 0, None] * items_t[0][None, :]).astype(F32)
    for k in range(1, items_t.shape[0]):
        acc = (acc + (Q[:, k, None] * items_t[k][None, :]).astype(F32)).astype(F32)
    return acc


def afm_topk_scores(A, w, n_user, n_item):
    """AFM.topk (AFM.py:209-246): un-normalised exp attention over user/context pairs and item x field
    pairs; score = (sum a*P / sum a) . w_pred + b[n]."""
    V = _f32(w["feature_embeddings"]); W = _f32(w["attention_W"])
    ab = _f32(w["attention_b"]).reshape(-1); ap = _f32(w["attention_p"]).reshape(-1)
    wp = _f32(w["prediction"]).reshape(-1)
    UF = np.concatenate([V[A[:, 0]][:, None, :], V[A[:, 2:]]], axis=1)          # :210-212 [C,uf,K]
    uf_n = UF.shape[1]
    pi = pair_index(uf_n)
    if pi:
        I = np.array([p[0] for p in pi]); J = np.array([p[1] for p in pi])
        uf = (UF[:, I] * UF[:, J]).astype(F32)                                  # :214-220
        z = np.maximum((uf @ W).astype(F32) + ab, F32(0))
        a_uf = np.exp((z * ap).astype(F32).sum(axis=2, dtype=F32)).astype(F32)  # :223-224 [C,P]
        UFwise = (uf * a_uf[:, :, None]).astype(F32).sum(axis=1, dtype=F32)     # :232
        a_sum = a_uf.sum(axis=1, dtype=F32)
    else:
        UFwise = np.zeros((A.shape[0], V.shape[1]), F32); a_sum = np.zeros(A.shape[0], F32)
    items = V[n_user:n_user + n_item]
    out = np.empty((A.shape[0], n_item), F32)
    bias = _f32(w["feature_bias"]).reshape(-1)[n_user:n_user + n_item]
    for c in range(A.shape[0]):                                                 # row loop bounds memory
        ufi = (UF[c][None, :, :] * items[:, None, :]).astype(F32)              # :227  [N,uf,K]
        zi = np.maximum((ufi @ W).astype(F32) + ab, F32(0))
        a_i = np.exp((zi * ap).astype(F32).sum(axis=2, dtype=F32)).astype(F32)  # :230  [N,uf]
        Iw = (ufi * a_i[:, :, None]).astype(F32).sum(axis=1, dtype=F32)         # :233
        s1 = (UFwise[c][None, :] + Iw).astype(F32)                              # :234
        wgt = (a_sum[c] + a_i.sum(axis=1, dtype=F32)).astype(F32)               # :235
        s2 = (s1 / wgt[:, None]).astype(F32)                                    # :236
        out[c] = ((s2 * wp).astype(F32).sum(axis=1, dtype=F32) + bias).astype(F32)   # :239-243
    return out


def hhfm_topk_scores(A, V, n_user, n_item, n_ctx, n_time, pools=(0, 0, 0)):
    """OUR.topk (OurModel7.py:229-295): A = [user, item, ctx..., time...]; hyb as in training."""
    Fea = A[:, 2:2 + n_ctx] if n_ctx > 0 else None
    Tim = A[:, 2 + n_ctx:2 + n_ctx + n_time] if n_time > 0 else None
    hyb, _ = hybrid_feature_seq(V, A[:, :2], Fea, Tim, pools)
    return dot_topk_scores(hyb, V, n_user, n_item)


def hybrid_feature_seq(V, Pos, Fea, Tim, pools):
    """hybrid_feature with the canonical sequential pooling order used by the top-N query builder."""
    V = _f32(V)

    def pool(E, mode):
        if mode == POOL_MAX:
            return E.max(axis=1)
        s = seq_sum(E, axis=1)
        return s if mode == POOL_SUM else (s / F32(E.shape[1])).astype(F32)

    stack = [V[Pos[:, 0]]]
    if Fea is not None and Fea.shape[1] > 0:
        stack.append(pool(V[Fea], pools[0]))
    if Tim is not None and Tim.shape[1] > 0:
        stack.append(pool(V[Tim], pools[1]))
    if len(stack) == 1:
        return stack[0], None
    return pool(np.stack(stack, axis=1), pools[2]).astype(F32), None


# ----------------------------------------------------------------------------------------------------
# AUC: FM.py:296-324
# ----------------------------------------------------------------------------------------------------
def auc_from_scores(pos_score, neg_score):
    """fraction of (positive, sampled negative) pairs with pos > neg (FM.py:320-324).
    pos_score [n], neg_score [n, 50]."""
    return float(np.mean((np.asarray(pos_score)[:, None] > np.asarray(neg_score)).reshape(-1)))
