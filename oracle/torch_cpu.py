"""PyTorch-CPU mirror of the reference's TF-1.x graphs, op for op  --  TEST INFRASTRUCTURE ONLY.

Purpose: (1) cross-check the analytic gradients of `oracle/hhfm_oracle.py` with torch.autograd;
(2) serve as the timed CPU baseline (`bench.py` cpu_baseline / `--impl reference`): TensorFlow cannot be
installed here, so the "reference CPU path" is this restatement run with all host threads.
It materialises the same intermediates the TF graph does ([B,F,K] gathers, [B,N,K] broadcasts) and lets
autograd produce the scatter-add the way TF's autodiff produces IndexedSlices / UnsortedSegmentSum.
Label every number produced with it "restatement, not TF".  The product never imports this file.
"""
from __future__ import annotations

import torch

POOL_SUM, POOL_MAX, POOL_MEAN = 0, 1, 2


def _pool(E, mode, dim=1):
    if mode == POOL_SUM:
        return E.sum(dim)
    if mode == POOL_MEAN:
        return E.mean(dim)
    return E.amax(dim)          # amax splits the gradient equally among ties, like tf.reduce_max


# ---- FM.py:99-126 ----------------------------------------------------------------------------------
def fm_out(X, V, b, b0):
    E = V[X]                                            # embedding_lookup  [B,F,K]
    S = E.sum(1, keepdim=True)
    fm = 0.5 * (S.square() - E.square().sum(1, keepdim=True))
    bil = fm.sum(1).sum(1, keepdim=True)
    fb = b[X].sum(1)
    return bil + fb + b0


def fm_loss(X, Y, V, b, b0, lamda):
    out = fm_out(X, V, b, b0)
    loss = 0.5 * (Y - out).square().sum()
    if lamda > 0:
        loss = loss + lamda * 0.5 * V.square().sum()
    return loss, out


# ---- MF.py:81-98 -----------------------------------------------------------------------------------
def mf_loss(X, Y, V, lamda):
    out = (V[X[:, 0]] * V[X[:, 1]]).sum(1, keepdim=True)
    loss = 0.5 * (Y - out).square().sum()
    if lamda > 0:
        loss = loss + lamda * 0.5 * V.square().sum()
    return loss, out


# ---- OurModel7.py:105-184 / BPR.py:76-88 -----------------------------------------------------------
def hybrid(V, Pos, Fea=None, Tim=None, pools=(0, 0, 0)):
    stack = [V[Pos[:, 0]]]
    if Fea is not None and Fea.shape[1] > 0:
        stack.append(_pool(V[Fea], pools[0]))
    if Tim is not None and Tim.shape[1] > 0:
        stack.append(_pool(V[Tim], pools[1]))
    if len(stack) == 1:
        return stack[0]
    return _pool(torch.stack(stack, 1), pools[2])


def pairrank_loss(V, Pos, Neg, Fea=None, Tim=None, pools=(0, 0, 0), lamda=0.0):
    hyb = hybrid(V, Pos, Fea, Tim, pools)
    pos = (hyb * V[Pos[:, 1]]).sum(1, keepdim=True)
    neg = (hyb.unsqueeze(1) * V[Neg]).sum(2, keepdim=True)
    mx = neg.amax(1)
    loss = -torch.log(torch.sigmoid(pos - mx)).sum()
    if lamda > 0:
        loss = loss + lamda * 0.5 * V.square().sum()
    return loss, pos, neg


# ---- AFM.py:103-148 --------------------------------------------------------------------------------
def afm_out(X, w):
    E = w["feature_embeddings"][X]
    F = X.shape[1]
    prods = [E[:, i, :] * E[:, j, :] for i in range(F) for j in range(i + 1, F)]
    P = torch.stack(prods).transpose(0, 1)                              # [B,P,K]
    K = P.shape[2]
    mul = (P.reshape(-1, K) @ w["attention_W"]).reshape(P.shape[0], P.shape[1], -1)
    s = (w["attention_p"] * torch.relu(mul + w["attention_b"])).sum(2, keepdim=True)
    a = torch.softmax(s, dim=1)
    afm = (a * P).sum(1)
    bil = (afm @ w["prediction"]).sum(1, keepdim=True)
    fb = w["feature_bias"][X].sum(1)
    return bil + fb + w["bias"]


def afm_loss(X, Y, w, lamda_attention):
    out = afm_out(X, w)
    loss = 0.5 * (Y - out).square().sum()
    if lamda_attention > 0:
        loss = loss + lamda_attention * 0.5 * w["attention_W"].square().sum()
    return loss, out


# ---- TF1 optimizers --------------------------------------------------------------------------------
@torch.no_grad()
def adagrad_(w, acc, g, lr):
    acc.add_(g * g)
    w.sub_(lr * g * acc.rsqrt())


# ---- top-N as the reference graph materialises it (FM.py:174-185, BPR.py:132-135) ------------------
@torch.no_grad()
def fm_topk(A, V, b, n_user, n_item, tp):
    user = V[A[:, 0]]
    item = V[n_user:n_user + n_item]
    feat = V[A[:, 2:]].sum(1)
    UF = user + feat
    IF = item.unsqueeze(0) + feat.unsqueeze(1)                          # [C,N,K] materialised like the reference
    score = (UF.unsqueeze(1) * IF).sum(2)
    bias = b[n_user:n_user + n_item].reshape(1, -1)
    return torch.topk(bias + score, tp, dim=1)


@torch.no_grad()
def dot_topk(Q, V, n_user, n_item, tp, materialise=True):
    item = V[n_user:n_user + n_item]
    if materialise:                                                     # OurModel7.py:294 broadcast-multiply-reduce
        score = (Q.unsqueeze(1) * item.unsqueeze(0)).sum(2)
    else:                                                               # BPR.py:134 matmul
        score = Q @ item.t()
    return torch.topk(score, tp, dim=1)


# ---- one full training step (forward, autodiff, TF1 Adagrad) ---------------------------------------
def hhfm_train_step(V, acc, Pos, Neg, Fea, Tim, pools, lamda, lr):
    V.requires_grad_(True)
    loss, _, _ = pairrank_loss(V, Pos, Neg, Fea, Tim, pools, lamda)
    (g,) = torch.autograd.grad(loss, V)
    V.requires_grad_(False)
    adagrad_(V, acc, g, lr)
    return float(loss.detach())


def fm_train_step(V, b, b0, accV, accb, accb0, X, Y, lamda, lr):
    for t in (V, b, b0):
        t.requires_grad_(True)
    loss, _ = fm_loss(X, Y, V, b, b0, lamda)
    gV, gb, gb0 = torch.autograd.grad(loss, (V, b, b0))
    for t in (V, b, b0):
        t.requires_grad_(False)
    adagrad_(V, accV, gV, lr)
    adagrad_(b, accb, gb, lr)
    adagrad_(b0, accb0, gb0, lr)
    return float(loss.detach())
