"""torch.distributed plumbing for the two places the path exchanges data:

  training   batch rows sharded across ranks; ONE all-reduce of the flat gradient arena per step
             ([gV | gbias | gb0 | loss partials], 1.4 MB at frappe shape), then every rank applies the identical update.
  evaluation item catalog sharded in contiguous id ranges; every rank scores all contexts against its items, keeps its
             local top-tp with GLOBAL ids, all-gather of [C, tp] (score, id), and a merge under the comparator
             (score desc, id asc) -- the lists are bit-identical to a single-GPU run.

`PeerArena` is the training exchange on one NVLink box: the gradient arenas are cudaMalloc'd, mapped into every peer
through CUDA IPC, and the optimizer kernel reads all ranks' arenas directly (csrc/p2p.cu) after a flag barrier -- no
all-reduce pass.  `allreduce_arena` (NCCL / gloo) remains the fallback and what the CPU tests exercise.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world(group=None):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def shard_range(n, rank, world_size):
    """Contiguous, balanced [lo, hi) of n units for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_arena(arena, group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(arena, op=dist.ReduceOp.SUM, group=group)
    return arena


def merge_topk(local_scores, local_ids, tp, group=None):
    """All-gather the per-shard candidate lists and keep the tp best per row under (score desc, id asc).
    local_scores f32 [C, tp_local], local_ids int [C, tp_local] (global ids; -1 = padding)."""
    rank, ws = world(group)
    if ws > 1:
        sc = [torch.empty_like(local_scores) for _ in range(ws)]
        ids = [torch.empty_like(local_ids) for _ in range(ws)]
        dist.all_gather(sc, local_scores.contiguous(), group=group)
        dist.all_gather(ids, local_ids.contiguous(), group=group)
        scores, idx = torch.cat(sc, dim=1), torch.cat(ids, dim=1)
    else:
        scores, idx = local_scores, local_ids
    if scores.is_cuda:
        from . import _lib
        from .engine import cur_stream, ptr
        C, n = scores.shape
        out_ids = torch.empty(C, tp, dtype=torch.int32, device=scores.device)
        out_sc = torch.empty(C, tp, dtype=torch.float32, device=scores.device)
        scores = scores.contiguous().float(); idx = idx.contiguous().to(torch.int32)
        # padding entries carry id -1: give them the lowest possible score so they sort last
        scores = torch.where(idx < 0, torch.full_like(scores, float("-inf")), scores)
        idx = torch.where(idx < 0, torch.full_like(idx, 2 ** 31 - 1), idx)
        _lib.call("hhfm_topn_select", ptr(scores), ptr(idx), None, C, n, n, tp, 0, ptr(out_sc), ptr(out_ids), cur_stream())
        out_ids = torch.where(out_ids == 2 ** 31 - 1, torch.full_like(out_ids, -1), out_ids)
        return out_ids, out_sc
    # host tensors (gloo tests / tiny merges): stable sort by id, then stable sort by descending score
    scores = torch.where(idx < 0, torch.full_like(scores, float("-inf")), scores)
    o1 = torch.argsort(idx, dim=1, stable=True)
    s1, i1 = torch.gather(scores, 1, o1), torch.gather(idx, 1, o1)
    s1 = s1 + 0.0                                                 # -0.0 -> +0.0 (top_k compares values)
    o2 = torch.argsort(s1, dim=1, descending=True, stable=True)
    return torch.gather(i1, 1, o2)[:, :tp], torch.gather(s1, 1, o2)[:, :tp]


def gather_rows(t, group=None):
    """Concatenate per-rank row blocks of possibly different length (rank order)."""
    rank, ws = world(group)
    if ws == 1:
        return t
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(outs, sizes)], dim=0)


class _RawDeviceBuffer:
    """Zero-copy view of library-owned device memory for torch.as_tensor (__cuda_array_interface__)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerArena:
    """Double-buffered EXPORT copy of the gradient arena, shared with the other ranks of the box over CUDA IPC + NVLink
    (csrc/p2p.cu).  The arena proper stays ordinary device memory; each step its gradient part is copied into bufs[cur].

    bufs[b]      this rank's export buffer b as a float32 tensor
    table(b, o)  host array with the address of float `o` of arena b on every rank (rank order)
    barrier()    stream-ordered cross-GPU barrier: every rank's scatter kernels of this step are complete and visible
    """

    FLAG_INTS = 64

    def __init__(self, n_floats, device, group=None):
        import ctypes as C
        from . import _lib
        self._lib = _lib
        self.rank, self.ws = world(group)
        if self.ws > 16:
            raise _lib.HhfmError("PeerArena: at most 16 ranks")
        self.device = device
        self.n = int(n_floats)
        nbytes = (self.n * 4 + 255) // 256 * 256
        self._own = []
        handles = []
        for _ in range(2):
            p = C.c_void_p()
            h = (C.c_ubyte * 64)()
            _lib.call("hhfm_p2p_alloc", nbytes, C.byref(p), C.cast(h, C.c_void_p))
            self._own.append(int(p.value))
            handles.append(bytes(h))
        p = C.c_void_p()
        h = (C.c_ubyte * 64)()
        _lib.call("hhfm_p2p_alloc", self.FLAG_INTS * 4, C.byref(p), C.cast(h, C.c_void_p))
        self._own_flags = int(p.value)
        handles.append(bytes(h))
        torch.cuda.synchronize(device)
        everyone = [None] * self.ws
        dist.all_gather_object(everyone, handles, group=group)
        self._opened = []
        self.base = [[0] * self.ws for _ in range(2)]
        self.flag_base = [0] * self.ws
        for r in range(self.ws):
            for k in range(3):
                if r == self.rank:
                    addr = (self._own + [self._own_flags])[k]
                else:
                    q = C.c_void_p()
                    hb = (C.c_ubyte * 64).from_buffer_copy(everyone[r][k])
                    _lib.call("hhfm_p2p_open", C.cast(hb, C.c_void_p), C.byref(q))
                    addr = int(q.value)
                    self._opened.append(addr)
                if k < 2:
                    self.base[k][r] = addr
                else:
                    self.flag_base[r] = addr
        self.bufs = [torch.as_tensor(_RawDeviceBuffer(self._own[b], (self.n,), "<f4"), device=device) for b in range(2)]
        self._flag_table = (C.c_int64 * self.ws)(*self.flag_base)
        self._tables = {}
        self.epoch = 0
        self.cur = 0
        dist.barrier(group=group)

    def table(self, b, offset_floats):
        import ctypes as C
        key = (b, int(offset_floats))
        t = self._tables.get(key)
        if t is None:
            t = (C.c_int64 * self.ws)(*[self.base[b][r] + 4 * int(offset_floats) for r in range(self.ws)])
            self._tables[key] = t
        return t

    def barrier(self):
        from .engine import cur_stream
        self.epoch += 1
        self._lib.call("hhfm_p2p_barrier", self._flag_table, self.rank, self.ws, self.epoch, cur_stream())

    def close(self):
        for a in self._opened:
            self._lib.call("hhfm_p2p_close", a)
        self._opened = []


def allreduce_metrics(codes, group=None):
    """HR / NDCG / reciprocal-rank over context rows sharded across ranks: every rank walks its rows
    (rank codes from `engine.metrics_walk`: -2 row dropped, -1 miss, n >= 0 hit at rank n), the three sums and the count
    of contributing rows are all-reduced (the count varies per rank: FM.py:336-357 drops rows), then divided.
    float64 sums: identical to np.average over the concatenated rows up to summation order."""
    import numpy as np
    c = np.asarray(codes).reshape(-1)
    hit = c >= 0
    n = c[hit].astype(np.float64)
    local = torch.tensor([float(hit.sum()), float((np.log(2.0) / np.log(n + 2.0)).sum()), float((1.0 / (n + 1.0)).sum()),
                          float((c != -2).sum())], dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        backend = dist.get_backend(group)
        t = local.cuda() if backend == "nccl" else local
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        local = t.cpu()
    cnt = float(local[3])
    if cnt == 0:
        return [float("nan")] * 3
    return [float(local[0]) / cnt, float(local[1]) / cnt, float(local[2]) / cnt]


def allgather_rows(ids, cols, group=None):
    """Coalesced-sparse exchange: every rank contributes `ids` [n] (int32, unique) and per-row payloads
    `cols` = list of tensors with leading dimension n; returns, for every rank r in rank order, (ids_r, [payload_r, ...]).
    The counts differ between ranks, so the lists travel padded to the longest one (one host sync to learn the counts).
    Pure torch.distributed plumbing: works on gloo (CPU tests) and NCCL alike."""
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    n = int(ids.shape[0])
    cnt = torch.tensor([n], dtype=torch.int64, device=ids.device)
    counts = [torch.zeros_like(cnt) for _ in range(ws)]
    dist.all_gather(counts, cnt, group=group)
    counts = [int(c.item()) for c in counts]
    maxc = max(counts)
    if maxc == 0:
        return [(ids[:0], [c[:0] for c in cols]) for _ in range(ws)]

    def gather(t, fill):
        pad = torch.full((maxc,) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=t.device)
        pad[:n] = t
        outs = [torch.empty_like(pad) for _ in range(ws)]
        dist.all_gather(outs, pad, group=group)
        return outs

    g_ids = gather(ids, -1)
    g_cols = [gather(c, 0) for c in cols]
    return [(g_ids[r][:counts[r]], [g[r][:counts[r]] for g in g_cols]) for r in range(ws)]
