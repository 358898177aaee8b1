"""Multi-GPU readout: one process per GPU; NVLink peer memory (or torch.distributed / NCCL) for the exchange.

Two ways the path shards:

1. independent sequences / objects -> plain data parallelism, no collective: every rank runs its own
   ``MemoryManager`` (``partition_sequences`` deals sequences to ranks; used by bench.py --gpus N).

2. one LVOS-scale long-term bank sharded along N (``ShardedLongTermReadout``, BASELINE.json configs[3]).  Rank r
   owns keys [lo_r, hi_r) AND the query rows [q_lo_r, q_hi_r); values are replicated.  Per frame and rank:

     select + push   fused similarity + LOCAL top-k over the rank's keys for all HW queries; the (score, global index)
                     list of every query goes straight to the rank that OWNS the query.
                       'peer' : the merge kernel stores into the owner's peer-mapped exchange buffer over NVLink and
                                raises one flag per owner (system-scope release) -- no collective, no barrier;
                       'nccl' : the lists are staged locally and exchanged by ONE all-to-all (NCCL; gloo in the CPU
                                protocol test), HW / world * k * 8 bytes per (source, owner) pair.
     exchange readout  waits for the source ranks' flags ('peer'), merges the `world` lists of each of its own queries
                     (global top-k is a subset of the union of local top-k's), softmax, sparse readout of its query
                     slice from the replicated value shadow.
     gather (optional)  every rank's slice is replicated on all ranks ('peer': stores into the peers' output buffers +
                     flags; 'nccl': all-gather).  Callers that consume the slice in place (a decoder sharded the same
                     way, or a host copy per rank) skip it: ``match(..., gather=False)``.

   Exchanged bytes per rank and frame: HW * k * 8 * (world-1)/world (245 KB per pair at LVOS / 8 ranks) instead of an
   all-gather's HW * k * 8 per source; the readout is not repeated on every rank.  The reference has no counterpart
   (single GPU, tools/runner.py:32); results equal the unsharded MemoryManager.match_memory, which the tests check.

The host logic is backend-agnostic so that the protocol is testable with gloo on CPU (tests inject a CPU backend built
on the oracle); the product backend is ``CudaBackend`` -- kernels of libvosmem.so only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

EXCH_K = 32          # entries per (source rank, query) exchange list (vosmem.h: VOSMEM_EXCH_K)
ENTRY = 8            # bytes per entry: fp32 score + int32 global key index


def shard_bounds(n: int, world: int, rank: int, align: int = 64) -> Tuple[int, int]:
    """Contiguous, tile-aligned split of [0, n) into `world` ranges; the last ranks may be empty."""
    per = -(-n // world)
    per = -(-per // align) * align
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def query_slices(hw: int, world: int, align: int = 16) -> int:
    """Queries owned by each rank (owner(q) = q // per); slice starts stay 64-byte aligned in fp32 rows."""
    per = -(-hw // world)
    return -(-per // align) * align


def partition_sequences(n_sequences: int, world: int, rank: int) -> List[int]:
    """Round-robin deal of independent sequences to ranks (data parallel, no collective)."""
    return list(range(rank, n_sequences, world))


@dataclass
class ShardProblem:
    """One object group of a sharded ``MemoryManager.match_memory`` call, as this rank sees it."""
    segments: list          # ops.KeySegment: this rank's part of [long-term | working] candidate keys (may be empty ranges)
    index_base: int         # local candidate i < seg0_len  -> global candidate i + index_base
    seg0_len: int
    index_base1: int        # local candidate i >= seg0_len -> global candidate i - seg0_len + index_base1
    values: list            # ops.ValueSegment on the GLOBAL candidate axis (banks are replicated)
    rows: int
    n_total: int
    share_ok: bool = True   # every rank scans at least one key (thresholds across ranks need all kernels running)


def plan_shard_ranges(ranges: Sequence[Tuple[int, int]], sizes: Sequence[int], world: int, rank: int
                      ) -> Tuple[List[Tuple[int, int]], int, int, int, bool]:
    """Candidate ranges of one object group, [begin_b, size_b) of every bank b (memory_manager.py:83,88-98: a group that
    entered late sees a suffix), cut down to what `rank` scans: bank b is dealt to the ranks in 64-key aligned blocks of
    its whole key axis (``shard_bounds(size_b, world, rank)``), so the blocks do not move while a group's suffix does.
    Returns (this rank's [lo, hi) per bank, index_base, seg0_len, index_base1, every rank has a key) with the GLOBAL
    candidate axis = the concatenated full ranges."""
    mine, counts = [], [0] * world
    for (begin, end), size in zip(ranges, sizes):
        assert end == size and 0 <= begin <= end
        for r in range(world):
            lo, hi = shard_bounds(size, world, r)
            lo, hi = max(lo, begin), max(hi, max(lo, begin))
            counts[r] += hi - lo
            if r == rank:
                mine.append((lo, hi))
    first = [0]
    for (begin, end) in ranges[:-1]:
        first.append(first[-1] + end - begin)
    index_base = mine[0][0] - ranges[0][0]
    seg0_len = mine[0][1] - mine[0][0]
    index_base1 = first[1] + mine[1][0] - ranges[1][0] if len(ranges) > 1 else 0
    return mine, index_base, (seg0_len if len(ranges) > 1 else -1), index_base1, all(c > 0 for c in counts)


class CudaBackend:
    """The product backend: every call is a libvosmem.so kernel on the current CUDA stream."""

    def __init__(self, device, value_dtype=torch.bfloat16):
        from . import _native, ops
        from .kv_memory_store import KeyValueMemoryStore
        self.ops, self.N, self.device = ops, _native, device
        self.store_cls, self.value_dtype = KeyValueMemoryStore, value_dtype
        self.keys = None
        self.values = None
        self.n_keys = 0

    def load_keys(self, key, shrinkage):
        self.keys = self.store_cls(count_usage=False, value_dtype=self.value_dtype)
        # the key bank carries no values of its own here (values are replicated separately)
        self.n_keys = key.shape[-1]
        if self.n_keys == 0:     # empty shard: the store holds one key that no candidate range ever covers
            key, shrinkage = torch.zeros((1, key.shape[1], 1)), torch.ones((1, 1, 1))
        self.keys.add(key.to(self.device), [], shrinkage.to(self.device), None, None)

    def load_values(self, value):
        """value: n_obj x CV x N (full bank, replicated)."""
        n_obj, cv, n = value.shape
        self.values = torch.empty((n, n_obj * cv), dtype=self.value_dtype, device=self.device)
        v = value.to(self.device).reshape(n_obj * cv, n)
        self.ops.pack_values(v, 0, n, self.values, 0)
        return n_obj * cv

    def new_buffer(self, nbytes: int) -> torch.Tensor:
        return torch.zeros(nbytes, dtype=torch.uint8, device=self.device)

    def workspace_bytes(self, ck: int, hw: int) -> int:
        return int(self.N.lib.vosmem_workspace_bytes(ck, hw, 0))

    def init_workspace(self, ws: torch.Tensor) -> None:
        self.N.check(self.N.lib.vosmem_workspace_init(ws.data_ptr(), ws.numel(), self.ops._stream()), 'vosmem_workspace_init')

    def select_push(self, qk, qe, top_k, index_base, per, world, rank, dst_ptrs, flag_ptrs, seq, ticket_ptr, send=None,
                    workspace=None, rank_pub_ptrs=None, segments=None, seg0_len=-1, index_base1=0):
        """Local selection over this rank's keys + push of every query's list to its owner (one C call, two launches).
        `send`: the tensor behind dst_ptrs in collective mode (unused here: the kernel takes the raw addresses).
        `workspace`: the engine's own (peer-mapped) scratch; `rank_pub_ptrs`: every rank's threshold summary array
        (thresholds shared across the ranks while the kernels run), or None.
        `segments`: this rank's candidate ranges as ops.KeySegment ([long-term shard | working-memory shard] of a
        sharded MemoryManager; None: the bank loaded with load_keys); local candidate i of them becomes global index
        i + index_base for i < seg0_len, i - seg0_len + index_base1 otherwise (seg0_len < 0: one base)."""
        N, ops = self.N, self.ops
        push = N.PushDesc()
        push.world, push.rank, push.per, push.index_base = world, rank, per, index_base
        push.seg0_len, push.index_base1 = seg0_len, index_base1
        for r in range(world):
            push.dst[r] = dst_ptrs[r]
            push.flag[r] = flag_ptrs[r] if flag_ptrs is not None else None
            push.rank_pub[r] = rank_pub_ptrs[r] if rank_pub_ptrs is not None else None
        push.seq, push.ticket = seq, ticket_ptr
        keep: list = []
        if segments is None:     # (an empty shard pushes empty lists: vosmem_select_push)
            segments = [self.keys.key_segment(0, self.n_keys)]
        sd = ops._select_desc(qk, qe, segments, top_k, 0, N.PATH_AUTO, keep, workspace=workspace)
        N.check(N.lib.vosmem_select_push(C.byref(sd), C.byref(push), ops._stream()), 'vosmem_select_push')

    def exchange_readout(self, lists_ptr, n_lists, list_stride, first_entry, flags_ptr, seq, status_ptr, n_q, top_k, rows,
                         n_total, out, lists=None, values=None):
        """out: rows x n_q view (row pitch = out.stride(0)) of this rank's query slice.  `lists`: the tensor behind
        lists_ptr in collective mode (unused here).  `values`: ops.ValueSegment list on the GLOBAL candidate axis
        (None: the replicated bank of load_values)."""
        N, ops = self.N, self.ops
        if values is None:
            values = [ops.ValueSegment(shadow=self.values, first=0, count=n_total, use_count=None)]
        rd = ops._readout_desc(n_q, top_k, rows, values, out, None)
        x = N.ExchangeDesc()
        x.lists, x.n_lists, x.list_stride, x.first_entry = lists_ptr, n_lists, list_stride, first_entry
        x.flags, x.seq, x.status = flags_ptr, seq, status_ptr
        N.check(N.lib.vosmem_exchange_readout(C.byref(rd), C.byref(x), ops._stream()), 'vosmem_exchange_readout')
        return out

    def push_slice(self, src, dst_ptrs, dst_ld, flag_ptrs, seq, ticket_ptr):
        N = self.N
        n = len(dst_ptrs)
        dp = (C.c_void_p * n)(*dst_ptrs)
        fp = (C.c_void_p * n)(*flag_ptrs)
        N.check(N.lib.vosmem_push_slice(src.data_ptr(), src.stride(0), src.shape[0], src.shape[1], dp, dst_ld, fp, n, seq,
                                        ticket_ptr, self.ops._stream()), 'vosmem_push_slice')

    def wait_flags(self, flags_ptr, mask, seq, status_ptr):
        self.N.check(self.N.lib.vosmem_wait_flags(flags_ptr, mask, seq, status_ptr, self.ops._stream()), 'vosmem_wait_flags')


class _Layout:
    """Byte layout of one rank's exchange buffer (identical on every rank: peers address it by offset)."""
    CONTROL = 4096          # list flags [2][16] u32 @0, output flags [2][16] u32 @128, tickets @256/@260, status @264

    def __init__(self, world: int, per: int, rows: int, hw: int, with_output: bool, ws_bytes: int = 0):
        self.world, self.per, self.hw = world, per, hw
        self.hw_pad = -(-hw // 128) * 128
        self.rank_pub0 = self.CONTROL                              # [world][hw_pad] x 8 B threshold summaries
        self.rank_pub_bytes = -(-world * self.hw_pad * 8 // 256) * 256
        self.ws0 = self.rank_pub0 + self.rank_pub_bytes            # the selection workspace (epochs in step across ranks)
        self.ws_bytes = -(-ws_bytes // 256) * 256
        self.list_bytes = per * EXCH_K * ENTRY                     # one source rank's lists for one owner
        self.slot_bytes = world * self.list_bytes
        self.lists0 = self.ws0 + self.ws_bytes
        self.out0 = self.lists0 + 2 * self.slot_bytes
        self.out_bytes = rows * hw * 4 if with_output else 0
        self.total = self.out0 + 2 * self.out_bytes

    def lists(self, slot: int, src: int = 0) -> int:
        return self.lists0 + slot * self.slot_bytes + src * self.list_bytes

    def list_flag(self, slot: int, src: int = 0) -> int:
        return (slot * 16 + src) * 4

    def out_flag(self, slot: int, src: int = 0) -> int:
        return 128 + (slot * 16 + src) * 4

    def out(self, slot: int) -> int:
        return self.out0 + slot * self.out_bytes

    ticket_push, ticket_slice, status = 256, 260, 264


class ShardedLongTermReadout:
    """Long-term memory bank sharded along N across `world` ranks (BASELINE.json configs[3])."""

    launches_per_match = 3      # select_tc (packs the query), merge_push, softmax_readout<exchange>  (+2 with gather)

    def __init__(self, config: dict, rank: int, world: int, device, backend=None, group=None):
        self.top_k = config['top_k']
        # 'peer': stores + flags over NVLink inside the kernels (CUDA backend only); 'nccl': one all-to-all of the lists
        self.exchange = str(config.get('vosmem_exchange', 'peer' if backend is None else 'nccl')).lower()
        assert self.exchange in ('nccl', 'peer')
        # 'n' (north_star): keys sharded, candidates exchanged, readout sharded over the queries.  'queries' (the control
        # of SURVEY section 8e): every rank keeps the whole bank and serves HW / world query rows end to end.
        self.shard = str(config.get('vosmem_shard', 'n')).lower()
        assert self.shard in ('n', 'queries')
        # thresholds shared across the ranks while the selection kernels run ('peer' mode only; see vosmem.h)
        self.share_thresholds = bool(config.get('vosmem_share_thresholds', True))
        self.rank, self.world, self.device, self.group = rank, world, device, group
        self.backend = backend if backend is not None else CudaBackend(device)
        self.n_total = 0
        self.rows = 0
        self.lo = self.hi = 0
        self.seq = 0
        self._buf = None         # (layout, local tensor, [base pointer of every rank])
        self._send = None

    # ------------------------------------------------------------------------------------------------
    def load_long_term(self, key, shrinkage, value) -> None:
        """key 1 x CK x N, shrinkage 1 x 1 x N, value n_obj x CV x N: the whole bank; this rank keeps keys [lo, hi)."""
        self.n_total = key.shape[-1]
        self.lo, self.hi = shard_bounds(self.n_total, self.world, self.rank) if self.shard == 'n' else (0, self.n_total)
        # (an empty shard still answers every query, with empty lists)
        self.backend.load_keys(key[:, :, self.lo:self.hi], shrinkage[:, :, self.lo:self.hi])
        self.rows = self.backend.load_values(value)

    def _buffers(self, hw: int, with_output: bool):
        """Exchange buffer of this rank, peer-mapped on every rank in 'peer' mode (allocated on first use: collective)."""
        per = query_slices(hw, self.world)
        if self._buf is not None and self._buf[0].per == per and self._buf[0].hw == hw and \
                (self._buf[0].out_bytes >= self.rows * hw * 4 or not with_output):
            return self._buf
        peer = self.exchange == 'peer' and self.world > 1
        ws_bytes = self.backend.workspace_bytes(64, hw) if peer and hasattr(self.backend, 'workspace_bytes') else 0
        lay = _Layout(self.world, per, self.rows, hw, with_output, ws_bytes)
        if peer:
            import torch.distributed._symmetric_memory as symm
            buf = symm.empty(lay.total, dtype=torch.uint8, device=self.device)
            buf.zero_()
            if ws_bytes:
                self.backend.init_workspace(buf[lay.ws0:lay.ws0 + ws_bytes])
            handle = symm.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
            bases = [int(p) for p in handle.buffer_ptrs]
            handle.barrier(channel=0)            # every rank's buffer is initialised before anyone pushes into it
            self._handle = handle
        else:
            buf = self.backend.new_buffer(lay.total)
            bases = [buf.data_ptr()] * self.world
        self._buf = (lay, buf, bases)
        return self._buf

    def query_range(self, hw: int) -> Tuple[int, int]:
        per = query_slices(hw, self.world)
        return min(hw, self.rank * per), min(hw, (self.rank + 1) * per)

    # ------------------------------------------------------------------------------------------------
    def match(self, query_key, selection, events=None, gather: bool = True, problem: Optional['ShardProblem'] = None
              ) -> torch.Tensor:
        """query_key / selection: 1 x CK x h x w.  gather=True -> rows x HW (the full readout on every rank);
        gather=False -> rows x (q_hi - q_lo), this rank's query slice (``query_range``).
        events (optional, for bench.py): [after local select + push, after exchange + readout, after gather].
        problem (optional): this rank's candidate ranges and the value segments of the call (a sharded
        MemoryManager passes its banks); None = the bank given to ``load_long_term``."""
        h, w = query_key.shape[-2:]
        hw = h * w
        qk = query_key.flatten(start_dim=2)[0]
        qe = selection.flatten(start_dim=2)[0] if selection is not None else None
        if problem is not None:
            assert self.shard == 'n'
            self.rows = max(self.rows, problem.rows)     # sizes the (symmetric) output buffers
        rows = problem.rows if problem is not None else self.rows
        sel_kw = dict(segments=problem.segments, seg0_len=problem.seg0_len, index_base1=problem.index_base1) \
            if problem is not None else {}
        rd_kw = dict(values=problem.values) if problem is not None else {}
        index_base = problem.index_base if problem is not None else self.lo
        n_total = problem.n_total if problem is not None else self.n_total
        if self.shard == 'queries':
            return self._match_query_sharded(qk, qe, hw, events, gather)
        be, k, G, r = self.backend, self.top_k, self.world, self.rank
        peer = self.exchange == 'peer' and G > 1
        lay, buf, bases = self._buffers(hw, with_output=gather and peer)
        per = lay.per
        q_lo, q_hi = self.query_range(hw)
        self.seq += 1
        seq, slot = self.seq, self.seq & 1
        me = bases[r]

        # ---- 1. local selection, lists pushed to the owners ----
        if peer:
            dst = [bases[d] + lay.lists(slot, r) for d in range(G)]
            flags = [bases[d] + lay.list_flag(slot, r) for d in range(G)]
            share = self.share_thresholds and lay.ws_bytes > 0 and qk.shape[0] == 64 and \
                (problem is None or problem.share_ok)
            be.select_push(qk, qe, k, index_base, per, G, r, dst, flags, seq, me + lay.ticket_push,
                           workspace=buf[lay.ws0:lay.ws0 + lay.ws_bytes] if lay.ws_bytes else None,
                           rank_pub_ptrs=[bases[d] + lay.rank_pub0 for d in range(G)] if share else None, **sel_kw)
        else:
            if self._send is None or self._send.numel() != lay.slot_bytes:
                self._send = be.new_buffer(lay.slot_bytes)
            send = self._send
            dst = [send.data_ptr() + d * lay.list_bytes for d in range(G)]
            be.select_push(qk, qe, k, index_base, per, G, r, dst, None, seq, me + lay.ticket_push, send=send, **sel_kw)
        if events is not None:
            events[0].record()

        # ---- 2. exchange (collective mode only) + merge of the `world` lists + softmax + readout of the own slice ----
        lists_ptr, flags_ptr, lists = me + lay.lists(slot), me + lay.list_flag(slot), None
        if not peer:
            if G > 1:
                lists = buf[lay.lists(slot):lay.lists(slot) + lay.slot_bytes]
                dist.all_to_all_single(lists, send, group=self.group)        # [owner][per][K] -> [source][per][K]
            else:
                lists, lists_ptr = send, send.data_ptr()
            flags_ptr = None
        n_q = q_hi - q_lo
        if gather and peer:
            full = buf[lay.out(slot):lay.out(slot) + rows * hw * 4].view(torch.float32).view(rows, hw)
        else:
            full = torch.empty((rows, hw if gather else max(n_q, 1)), dtype=torch.float32, device=qk.device)
        mine = full[:, q_lo:q_hi] if gather else full[:, :n_q]
        if n_q > 0:
            be.exchange_readout(lists_ptr, G, per * EXCH_K, 0, flags_ptr, seq, me + lay.status, n_q, k, rows,
                                n_total, mine, lists=lists, **rd_kw)
        if events is not None:
            events[1].record()
        if not gather:
            if events is not None:
                events[2].record()
            return mine

        # ---- 3. replicate the slices ----
        if G > 1 and peer:
            if n_q > 0:
                others = [d for d in range(G) if d != r]
                be.push_slice(mine, [bases[d] + lay.out(slot) + q_lo * 4 for d in others], hw,
                              [bases[d] + lay.out_flag(slot, r) for d in others], seq, me + lay.ticket_slice)
            mask = sum(1 << s for s in range(G) if s != r and min(hw, s * per) < min(hw, (s + 1) * per))
            if mask:                # ranks whose query slice is empty push nothing
                be.wait_flags(me + lay.out_flag(slot), mask, seq, me + lay.status)
        elif G > 1:
            part = torch.zeros((rows, per), dtype=torch.float32, device=qk.device)
            part[:, :n_q] = mine
            gathered = self._all_gather(part)                                   # world x rows x per
            full = gathered.permute(1, 0, 2).reshape(rows, G * per)[:, :hw]
        if events is not None:
            events[2].record()
        return full

    def check_status(self) -> None:
        """Raise if a flag wait timed out on the device (synchronises)."""
        if self._buf is None:
            return
        lay, buf, _ = self._buf
        if int(buf[lay.status:lay.status + 4].view(torch.int32).item()) != 0:
            raise RuntimeError('sharded readout: a wait for another rank\'s flag timed out')

    def _all_gather(self, t: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return t.unsqueeze(0)
        t = t.contiguous()
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return out.view((self.world,) + tuple(t.shape))

    def _match_query_sharded(self, qk, qe, hw, events, gather):
        """Control: query rows [q_lo, q_hi) of this rank against the whole bank (one all-gather of the slices if asked)."""
        be, k = self.backend, self.top_k
        per = query_slices(hw, self.world)
        q_lo, q_hi = self.query_range(hw)
        n_q = q_hi - q_lo
        part = torch.zeros((self.rows, per), dtype=torch.float32, device=qk.device)
        if n_q > 0:
            sl_k = qk[:, q_lo:q_hi].contiguous()
            sl_e = qe[:, q_lo:q_hi].contiguous() if qe is not None else None
            if self._send is None or self._send.numel() != n_q * EXCH_K * ENTRY + 64:
                self._send = be.new_buffer(n_q * EXCH_K * ENTRY + 64)
            base = self._send.data_ptr()
            be.select_push(sl_k, sl_e, k, 0, n_q, 1, 0, [base], None, 1, base + n_q * EXCH_K * ENTRY, send=self._send)
            if events is not None:
                events[0].record()
            be.exchange_readout(base, 1, n_q * EXCH_K, 0, None, 1, None, n_q, k, self.rows, self.n_total, part[:, :n_q],
                                lists=self._send)
        elif events is not None:
            events[0].record()
        if events is not None:
            events[1].record()
        if not gather or self.world == 1:
            if events is not None:
                events[2].record()
            return part[:, :n_q] if not gather else part[:, :hw]
        gathered = self._all_gather(part)                               # world x rows x per
        if events is not None:
            events[2].record()
        return gathered.permute(1, 0, 2).reshape(self.rows, self.world * per)[:, :hw]
