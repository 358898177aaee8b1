"""Drop-in for the reference's ``model.memory_util`` (tracker/model/memory_util.py).

Same names, argument meaning and return shapes as the reference; every function runs a kernel of
libvosmem.so on the tensors' CUDA device (CPU tensors are rejected -- there is no fallback).

These are the *dense* twins: they materialise the B x N x HW matrix like the reference does, and are
used by memory consolidation and by ``XMem.read_memory`` (training).  The per-frame hot path
(``MemoryManager.match_memory``) does not go through them; it calls the fused kernels.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


def get_similarity(mk, ms, qk, qe):
    """memory_util.py:7-39.  mk: B x CK x [N], ms: B x 1 x [N] | None, qk/qe: B x CK x [HW] -> B x N x HW."""
    B, CK = mk.shape[:2]
    mk = mk.flatten(start_dim=2)
    ms = ms.flatten(start_dim=1) if ms is not None else None
    qk = qk.flatten(start_dim=2)
    qe = qe.flatten(start_dim=2) if qe is not None else None
    outs = [ops.similarity_dense(mk[b], ms[b] if ms is not None else None, qk[b], qe[b] if qe is not None else None)
            for b in range(B)]
    return outs[0].unsqueeze(0) if B == 1 else torch.stack(outs, 0)


def do_softmax(similarity, top_k: Optional[int] = None, inplace=False, return_usage=False):
    """memory_util.py:41-65.  similarity: B x N x [HW] -> affinity B x N x HW [, usage B x N]."""
    if similarity.dim() != 3:
        raise RuntimeError(f'do_softmax expects B x N x HW, got {tuple(similarity.shape)}')
    B = similarity.shape[0]
    if top_k is not None and top_k > similarity.shape[1]:
        # torch.topk in the reference raises the same way (memory_util.py:46)
        raise RuntimeError(f'selected index k out of range (top_k={top_k}, N={similarity.shape[1]})')
    affs, usages = [], []
    for b in range(B):
        # the reference only honours `inplace` on the top-k branch (memory_util.py:50-54)
        a, u = ops.softmax_dense(similarity[b], top_k, inplace and top_k is not None, return_usage)
        affs.append(a)
        usages.append(u)
    if inplace and top_k is not None and all(a.data_ptr() == similarity[b].data_ptr() for b, a in enumerate(affs)):
        affinity = similarity
    else:
        affinity = affs[0].unsqueeze(0) if B == 1 else torch.stack(affs, 0)
    if return_usage:
        usage = usages[0].unsqueeze(0) if B == 1 else torch.stack(usages, 0)
        return affinity, usage
    return affinity


def get_affinity(mk, ms, qk, qe):
    """memory_util.py:67-71."""
    return do_softmax(get_similarity(mk, ms, qk, qe))


def readout(affinity, mv):
    """memory_util.py:73-80.  affinity B x THW x HW, mv B x CV x T x H x W -> B x CV x H x W."""
    B, CV, T, H, W = mv.shape
    mo = mv.reshape(B, CV, T * H * W)
    outs = [ops.readout_dense(mo[b], affinity[b]) for b in range(B)]
    mem = outs[0].unsqueeze(0) if B == 1 else torch.stack(outs, 0)
    return mem.view(B, CV, H, W)


__all__ = ['get_similarity', 'do_softmax', 'get_affinity', 'readout']
