"""Multi-GPU readout: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch) for the exchange.

Two ways the path shards:

1. independent sequences / objects -> plain data parallelism, no collective: every rank runs its own
   ``MemoryManager`` (``partition_sequences`` deals sequences to ranks; used by bench.py --gpus N).

2. one LVOS-scale long-term bank sharded along N (``ShardedLongTermReadout``):
     rank r owns keys [lo_r, hi_r) (packed image + fp32 keys) -> fused similarity + LOCAL top-k
     all-gather of the (score, global index) candidates: HW * k * 8 bytes per rank
     every rank merges the G candidate lists -> the same global top-k (vosmem_merge_topk)
     softmax + readout on every rank from a replicated value shadow (no second collective and no reduction
     anywhere: global top-k is a subset of the union of local top-k's, and the softmax needs only the k
     merged scores).
   The reference has no counterpart (single GPU, tools/runner.py:32); results equal the unsharded
   MemoryManager.match_memory by construction, which the tests check.

The host logic is backend-agnostic so that the protocol is testable with gloo on CPU (tests inject a
CPU backend built on the oracle); the product backend is ``CudaBackend`` -- kernels of libvosmem.so only.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int, align: int = 64) -> Tuple[int, int]:
    """Contiguous, tile-aligned split of [0, n) into `world` ranges; the last ranks may be empty."""
    per = -(-n // world)
    per = -(-per // align) * align
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def partition_sequences(n_sequences: int, world: int, rank: int) -> List[int]:
    """Round-robin deal of independent sequences to ranks (data parallel, no collective)."""
    return list(range(rank, n_sequences, world))


class CudaBackend:
    """The product backend: every call is a libvosmem.so kernel on the current CUDA stream."""

    def __init__(self, device, value_dtype=torch.bfloat16):
        from . import ops
        from .kv_memory_store import KeyValueMemoryStore
        self.ops, self.device = ops, device
        self.store_cls, self.value_dtype = KeyValueMemoryStore, value_dtype
        self.keys = None
        self.values = None

    def load_keys(self, key, shrinkage):
        self.keys = self.store_cls(count_usage=False, value_dtype=self.value_dtype)
        n = key.shape[-1]
        # the key bank carries no values of its own here (values are replicated separately)
        self.keys.add(key.to(self.device), [], shrinkage.to(self.device), None, None)
        return n

    def load_values(self, value):
        """value: n_obj x CV x N (full bank, replicated)."""
        n_obj, cv, n = value.shape
        self.values = torch.empty((n, n_obj * cv), dtype=self.value_dtype, device=self.device)
        v = value.to(self.device).reshape(n_obj * cv, n)
        self.ops.pack_values(v, 0, n, self.values, 0)
        return n_obj * cv

    def select(self, qk, qe, top_k, index_base, out=None):
        n = self.keys.size
        return self.ops.select_topk(qk, qe, [self.keys.key_segment(0, n)], top_k, index_base=index_base, out=out)

    def merge_ptrs(self, score_ptrs, index_ptrs, hw, top_k):
        return self.ops.merge_topk_ptrs(score_ptrs, index_ptrs, hw, top_k, self.device)

    def merge(self, scores, indices):
        return self.ops.merge_topk(scores, indices)

    def readout(self, score, index, rows, n_total, out):
        seg = self.ops.ValueSegment(shadow=self.values, first=0, count=n_total, use_count=None)
        return self.ops.softmax_readout(score, index, [seg], rows, out=out)


class PeerExchange:
    """Candidate exchange without a collective: every rank writes its local top-k lists into a symmetric (peer-mapped)
    buffer, one device-side barrier tells the ranks that all lists are in place, and the merge kernel loads the other
    ranks' lists over NVLink itself (vosmem_merge_topk_ptrs).  Two slots alternate between frames, so the single barrier
    per frame also protects the slot that is overwritten two frames later."""

    def __init__(self, hw: int, top_k: int, device, group):
        import torch.distributed._symmetric_memory as symm
        self.hw, self.k = hw, top_k
        self.score_bytes = hw * top_k * 4
        self.slot_bytes = hw * top_k * 12                      # fp32 scores + int64 global indices
        self.buf = symm.empty(2 * self.slot_bytes, dtype=torch.uint8, device=device)
        self.handle = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.peer_base = [int(p) for p in self.handle.buffer_ptrs]
        self.frame = 0

    def slot_views(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(score, index) views of this frame's slot in the local buffer: the selection writes straight into them."""
        b = (self.frame & 1) * self.slot_bytes
        score = self.buf[b:b + self.score_bytes].view(torch.float32).view(self.hw, self.k)
        index = self.buf[b + self.score_bytes:b + self.slot_bytes].view(torch.int64).view(self.hw, self.k)
        return score, index

    def peer_lists(self) -> Tuple[List[int], List[int]]:
        b = (self.frame & 1) * self.slot_bytes
        return [p + b for p in self.peer_base], [p + b + self.score_bytes for p in self.peer_base]

    def barrier_and_advance(self):
        self.handle.barrier(channel=0)
        self.frame += 1


class ShardedLongTermReadout:
    """Long-term memory bank sharded along N across `world` ranks (BASELINE.json configs[3])."""

    launches_per_match = 4  # select_tc (packs the query), merge_splits, merge_lists, softmax_readout

    def __init__(self, config: dict, rank: int, world: int, device, backend=None, group=None):
        self.top_k = config['top_k']
        # 'nccl': one all-gather of the packed candidates; 'peer': symmetric-memory buffers + NVLink loads inside the
        # merge kernel (CUDA backend only)
        self.exchange = str(config.get('vosmem_exchange', 'nccl')).lower()
        assert self.exchange in ('nccl', 'peer')
        self._peer = None
        # 'n' (north_star): the key axis is sharded, candidates are exchanged.  'queries' (the control of SURVEY section
        # 8e): every rank keeps the whole bank and serves HW / world query rows end to end; the only collective is the
        # all-gather of the rows x HW readout slices.
        self.shard = str(config.get('vosmem_shard', 'n')).lower()
        assert self.shard in ('n', 'queries')
        self.rank, self.world, self.device, self.group = rank, world, device, group
        self.backend = backend if backend is not None else CudaBackend(device)
        self.n_total = 0
        self.rows = 0
        self.lo = self.hi = 0

    def load_long_term(self, key, shrinkage, value) -> None:
        """key 1 x CK x N, shrinkage 1 x 1 x N, value n_obj x CV x N: the whole bank; this rank keeps keys [lo, hi)."""
        self.n_total = key.shape[-1]
        self.lo, self.hi = shard_bounds(self.n_total, self.world, self.rank) if self.shard == 'n' else (0, self.n_total)
        if self.hi > self.lo:
            self.backend.load_keys(key[:, :, self.lo:self.hi], shrinkage[:, :, self.lo:self.hi])
        self.rows = self.backend.load_values(value)

    def _all_gather(self, t: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return t.unsqueeze(0)
        t = t.contiguous()
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return out.view((self.world,) + tuple(t.shape))

    def match(self, query_key, selection, events=None) -> torch.Tensor:
        """query_key / selection: 1 x CK x h x w  ->  rows x HW (full readout on every rank).
        events (optional, for bench.py): [after local select, after exchange + merge, after readout]."""
        h, w = query_key.shape[-2:]
        hw = h * w
        qk = query_key.flatten(start_dim=2)[0]
        qe = selection.flatten(start_dim=2)[0] if selection is not None else None
        k = self.top_k
        if self.shard == 'queries':
            return self._match_query_sharded(qk, qe, hw, events)
        peer = None
        if self.exchange == 'peer' and self.world > 1 and hasattr(self.backend, 'merge_ptrs'):
            if self._peer is None or self._peer.hw != hw:
                self._peer = PeerExchange(hw, k, qk.device, self.group)      # collective: every rank gets here together
            peer = self._peer
        out = peer.slot_views() if peer is not None else None
        if self.hi > self.lo:
            score, index = self.backend.select(qk, qe, k, self.lo, out) if out is not None else \
                self.backend.select(qk, qe, k, self.lo)
        elif out is not None:   # empty shard: contributes no candidates
            score, index = out[0].fill_(float('-inf')), out[1].fill_(-1)
        else:
            score = torch.full((hw, k), float('-inf'), dtype=torch.float32, device=qk.device)
            index = torch.full((hw, k), -1, dtype=torch.int64, device=qk.device)
        if events is not None:
            events[0].record()
        if peer is not None:
            # ---- exchange fused into the merge: barrier, then the merge kernel reads every rank's list in place ----
            score_ptrs, index_ptrs = peer.peer_lists()
            peer.barrier_and_advance()
            g_score, g_index = self.backend.merge_ptrs(score_ptrs, index_ptrs, hw, k)
            all_s = None
        # ---- the one exchange step: all-gather of the local top-k candidates (one message: fp32 score bits and
        #      the global index as int32 pairs, HW * k * 8 bytes per rank) ----
        elif self.world == 1:
            all_s, all_i = score.unsqueeze(0), index.unsqueeze(0)
        else:
            packed = torch.stack((score.view(torch.int32), index.to(torch.int32)))     # 2 x HW x k
            gathered = self._all_gather(packed)                                          # world x 2 x HW x k
            all_s = gathered[:, 0].contiguous().view(torch.float32)
            all_i = gathered[:, 1].to(torch.int64)
        if all_s is not None:
            g_score, g_index = self.backend.merge(all_s, all_i)            # identical on every rank
        if events is not None:
            events[1].record()
        # ---- readout: values are replicated, so every rank reads out all query rows itself.  (Slicing the query
        #      rows over the ranks and all-gathering rows x HW fp32 afterwards was measured slower at 2 GPUs: 24 us
        #      for the half readout + 56 us for the 16.7 MB gather against 35 us for the whole readout.) ----
        out = torch.empty((self.rows, hw), dtype=torch.float32, device=qk.device)
        self.backend.readout(g_score, g_index, self.rows, self.n_total, out=out)
        if events is not None:
            events[2].record()
        return out

    def _match_query_sharded(self, qk, qe, hw, events):
        """Query rows [qlo, qhi) of this rank against the whole bank, then one all-gather of the readout slices."""
        per = -(-hw // self.world)
        per = -(-per // 16) * 16                                   # slice starts stay 64-byte aligned
        qlo, qhi = min(hw, self.rank * per), min(hw, (self.rank + 1) * per)
        part = torch.zeros((self.rows, per), dtype=torch.float32, device=qk.device)
        if qhi > qlo:
            sl_k = qk[:, qlo:qhi].contiguous()
            sl_e = qe[:, qlo:qhi].contiguous() if qe is not None else None
            score, index = self.backend.select(sl_k, sl_e, self.top_k, 0)
            if events is not None:
                events[0].record()
                events[1].record()
            self.backend.readout(score, index, self.rows, self.n_total, out=part[:, :qhi - qlo])
        elif events is not None:
            events[0].record()
            events[1].record()
        if events is not None:
            events[2].record()
        if self.world == 1:
            return part[:, :hw]
        gathered = self._all_gather(part)                               # world x rows x per
        return gathered.permute(1, 0, 2).reshape(self.rows, self.world * per)[:, :hw]
