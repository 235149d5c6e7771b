"""torch.autograd.Function wrappers over the C ABI, for composing the hot-path kernels with other torch code.

The model classes in `models.py` use the fused forward+backward entry points directly (one pass over the batch);
these Functions expose the same kernels as differentiable ops:

    out = fm_interaction(idx, V, bias, b0)              # FM.py:99-120      -> [B]
    pos, neg = pairrank_scores(idx, V, n_ctx, n_time, n_neg, pools)   # OurModel7.py:105-172 -> [B], [B,NG]

`idx` is an int32 CUDA tensor ([B,F] for FM; packed records [B,stride] for pairrank, see hhfm_sm100.h).
Gradients w.r.t. the table are returned dense ([M,K], accumulated with the sort-free vector reductions).
"""
from __future__ import annotations

import torch

from . import _lib
from .engine import cur_stream, ptr


class _FMInteraction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, V, bias, b0, interaction):
        if not (idx.is_cuda and V.is_cuda):
            raise _lib.HhfmError("fm_interaction: CUDA tensors required (no CPU fallback)")
        idx = idx.contiguous()
        V = V.contiguous()
        B, F = idx.shape
        out = torch.empty(B, dtype=torch.float32, device=V.device)
        _lib.call("hhfm_fm_fwd", None, ptr(idx), None, B, F, ptr(V), ptr(bias), ptr(b0), V.shape[0], V.shape[1],
                  interaction, ptr(out), cur_stream())
        ctx.save_for_backward(idx, V)
        ctx.has_bias = bias is not None
        ctx.has_b0 = b0 is not None
        ctx.bias_shape = bias.shape if bias is not None else None
        ctx.interaction = interaction
        return out

    @staticmethod
    def backward(ctx, gout):
        idx, V = ctx.saved_tensors
        B, F = idx.shape
        gout = gout.contiguous().float()
        gV = torch.zeros_like(V)
        gb = torch.zeros(V.shape[0], dtype=torch.float32, device=V.device) if ctx.has_bias else None
        gb0 = torch.zeros(1, dtype=torch.float32, device=V.device) if ctx.has_b0 else None
        _lib.call("hhfm_fm_bwd", None, ptr(idx), None, B, F, ptr(V), V.shape[0], V.shape[1], ctx.interaction, ptr(gout),
                  ptr(gV), ptr(gb), ptr(gb0), 0, cur_stream())
        return (None, gV, gb.view(ctx.bias_shape) if gb is not None else None, gb0, None)


def fm_interaction(idx, V, bias=None, b0=None, interaction=0):
    if b0 is not None:
        b0 = b0.reshape(1)          # differentiable view: the [1] gradient flows back to a 0-dim scalar
    return _FMInteraction.apply(idx, V, bias, b0, interaction)


class _PairRankScores(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, V, n_ctx, n_time, n_neg, pools):
        if not (idx.is_cuda and V.is_cuda):
            raise _lib.HhfmError("pairrank_scores: CUDA tensors required (no CPU fallback)")
        idx = idx.contiguous()
        V = V.contiguous()
        B, stride = idx.shape
        pos = torch.empty(B, dtype=torch.float32, device=V.device)
        neg = torch.empty(B, n_neg, dtype=torch.float32, device=V.device) if n_neg > 0 else None
        _lib.call("hhfm_pairrank_fwd", ptr(idx), B, stride, n_ctx, n_time, n_neg, pools[0], pools[1], pools[2], ptr(V),
                  V.shape[0], V.shape[1], ptr(pos), ptr(neg), cur_stream())
        ctx.save_for_backward(idx, V)
        ctx.cfg = (n_ctx, n_time, n_neg, tuple(pools))
        if neg is None:
            neg = torch.empty(B, 0, dtype=torch.float32, device=V.device)
        return pos, neg

    @staticmethod
    def backward(ctx, dpos, dneg):
        idx, V = ctx.saved_tensors
        n_ctx, n_time, n_neg, pools = ctx.cfg
        B, stride = idx.shape
        dpos = torch.zeros(B, device=V.device) if dpos is None else dpos.contiguous().float()
        dneg = dneg.contiguous().float() if (dneg is not None and n_neg > 0) else None
        gV = torch.zeros_like(V)
        _lib.call("hhfm_pairrank_bwd", ptr(idx), B, stride, n_ctx, n_time, n_neg, pools[0], pools[1], pools[2], ptr(V),
                  V.shape[0], V.shape[1], ptr(dpos), ptr(dneg), ptr(gV), 0, cur_stream())
        return None, gV, None, None, None, None


def pairrank_scores(idx, V, n_ctx, n_time, n_neg, pools=(0, 0, 0)):
    return _PairRankScores.apply(idx, V, n_ctx, n_time, n_neg, tuple(pools))


def scatter_add_rows(dst, rows, src):
    """dst[rows[i], :] += src[i, :] with the sort-free vector reduction kernel (UnsortedSegmentSum)."""
    rows = rows.contiguous()
    src = src.contiguous()
    K = 1 if src.dim() == 1 else src.shape[1]
    _lib.call("hhfm_scatter_add_rows", ptr(rows), ptr(src), rows.numel(), K, ptr(dst), dst.shape[0], cur_stream())
    return dst
