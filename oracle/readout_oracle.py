"""CPU restatement of VOS-E-SAM / XMem's space-time memory readout.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Parity status: PINNED.
The reference ships no tests or golden vectors for this path, so
the oracle is pinned against outputs of the reference itself: ``gen_golden.py``
imports the unmodified reference from ``/root/reference`` in the authoring
container, runs it on seeded inputs and commits the results under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function below
against those vectors.

Everything is written against plain ``torch`` CPU tensors because the reference
path *is* torch tensor algebra; the op sequence (two dense products, seven
elementwise passes over the N x HW matrix, ``topk``, ``scatter``, a dense
``values @ affinity``) is kept so that timing this file on the host cores is a
fair stand-in for timing the reference (``bench.py`` ``cpu_baseline.kind ==
"port"``).  All functions are dtype-agnostic: feed ``float64`` tensors to get
the fp64 scores that the parity rule's k / k+1 gap is computed from.

Reference lines restated (relative to /root/reference):
  tracker/model/memory_util.py:7-39      -> ``anisotropic_l2``
  tracker/model/memory_util.py:41-65     -> ``topk_affinity`` / ``dense_affinity``
  tracker/model/memory_util.py:73-80     -> ``readout_5d``
  tracker/inference/memory_manager.py:57-150 -> ``match``
  tracker/inference/kv_memory_store.py:36-99,135-214 -> ``Bank``
  tracker/inference/memory_manager.py:152-286 -> ``Readout.add_memory`` etc.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


# ---------------------------------------------------------------------------
# memory_util.py
# ---------------------------------------------------------------------------
def anisotropic_l2(mem_key: Tensor, mem_shrink: Optional[Tensor], q_key: Tensor,
                   q_sel: Optional[Tensor]) -> Tensor:
    """Similarity S[b, n, q] (memory_util.py:7-39).

    With a selection term e and shrinkage s this is
        S[n,q] = s[n]/sqrt(CK) * ( -sum_c k[c,n]^2 e[c,q]
                                   + 2 sum_c k[c,n] q[c,q] e[c,q]
                                   -   sum_c e[c,q] q[c,q]^2 )
    i.e. -s * e-weighted ||k - q||^2 / sqrt(CK) (memory_util.py:20-27).  Without
    a selection term the query-norm term is dropped (memory_util.py:28-32);
    without shrinkage only the 1/sqrt(CK) scale remains (memory_util.py:34-37).
    """
    ck = mem_key.shape[1]
    k = mem_key.reshape(mem_key.shape[0], ck, -1)          # B, CK, N
    q = q_key.reshape(q_key.shape[0], ck, -1)              # B, CK, HW
    kt = k.transpose(1, 2)                                 # B, N, CK
    if q_sel is not None:
        e = q_sel.reshape(q_sel.shape[0], ck, -1)
        sq_term = torch.matmul(kt * kt, e)                 # memory_util.py:24
        cross = torch.matmul(kt, q * e) * 2                # memory_util.py:25
        q_norm = (e * q * q).sum(dim=1, keepdim=True)      # memory_util.py:26
        s = cross - sq_term - q_norm                       # memory_util.py:27
    else:
        sq_term = (k * k).sum(dim=1).unsqueeze(2)          # memory_util.py:30
        cross = torch.matmul(kt, q) * 2                    # memory_util.py:31
        s = cross - sq_term                                # memory_util.py:32
    if mem_shrink is not None:
        shr = mem_shrink.reshape(mem_shrink.shape[0], -1).unsqueeze(2)
        s = s * shr / math.sqrt(ck)                        # memory_util.py:35
    else:
        s = s / math.sqrt(ck)                              # memory_util.py:37
    return s


def topk_select(sim: Tensor, k: int) -> Tuple[Tensor, Tensor, Tensor]:
    """Top-k over the memory axis + softmax of the survivors (memory_util.py:46-49).

    Returns (indices[B,k,HW] int64, weights[B,k,HW], values[B,k,HW]).  The
    reference does *not* subtract the maximum on this branch (memory_util.py:48).
    """
    vals, idx = torch.topk(sim, k=k, dim=1)
    w = vals.exp()
    w = w / w.sum(dim=1, keepdim=True)
    return idx, w, vals


def topk_affinity(sim: Tensor, k: int, want_usage: bool = False):
    """Dense affinity with exactly k non-zeros per query column (memory_util.py:45-54,62-63)."""
    idx, w, _ = topk_select(sim, k)
    aff = torch.zeros_like(sim).scatter_(1, idx, w)
    if want_usage:
        return aff, aff.sum(dim=2)
    return aff


def dense_affinity(sim: Tensor, want_usage: bool = False):
    """Max-subtracted softmax over the memory axis (memory_util.py:55-60)."""
    top = sim.max(dim=1, keepdim=True).values
    w = (sim - top).exp()
    aff = w / w.sum(dim=1, keepdim=True)
    if want_usage:
        return aff, aff.sum(dim=2)
    return aff


def affinity(sim: Tensor, top_k: Optional[int] = None, want_usage: bool = False):
    """Dispatcher with the reference's ``do_softmax`` semantics (memory_util.py:41-65)."""
    if top_k is None:
        return dense_affinity(sim, want_usage)
    return topk_affinity(sim, top_k, want_usage)


def readout_5d(aff: Tensor, mem_value: Tensor) -> Tensor:
    """Training-time twin: values B x CV x T x H x W times affinity (memory_util.py:73-80)."""
    b, cv, t, h, w = mem_value.shape
    flat = mem_value.reshape(b, cv, t * h * w)
    return torch.bmm(flat, aff).reshape(b, cv, h, w)


# ---------------------------------------------------------------------------
# kv_memory_store.py
# ---------------------------------------------------------------------------
class Bank:
    """Key / shrinkage / selection / per-group values + usage counters.

    Restates ``KeyValueMemoryStore`` (kv_memory_store.py:4-214).  Tensors keep
    the reference layout: key 1xCKxN, shrinkage 1x1xN, selection 1xCKxN,
    values[g] n_g x CV x N_g with N contiguous; later groups hold only a suffix
    of the key timeline (kv_memory_store.py:10-16).
    """

    def __init__(self, count_usage: bool):
        self.count_usage = count_usage
        self.key: Optional[Tensor] = None
        self.shrinkage: Optional[Tensor] = None
        self.selection: Optional[Tensor] = None
        self.values: List[Tensor] = []
        self.groups: List[List[int]] = []
        self.seen: List[int] = []
        self.use_count: Optional[Tensor] = None
        self.life_count: Optional[Tensor] = None

    # -- sizes -------------------------------------------------------------
    @property
    def size(self) -> int:
        return 0 if self.key is None else int(self.key.shape[-1])

    @property
    def num_groups(self) -> int:
        return len(self.values)

    def engaged(self) -> bool:
        return self.key is not None

    def group_len(self, gi: int) -> int:
        return int(self.values[gi].shape[2])

    # -- growth (kv_memory_store.py:36-90) -----------------------------------
    def append(self, key: Tensor, value, shrinkage: Optional[Tensor], selection: Optional[Tensor],
               objects: Optional[Sequence[int]]) -> None:
        n_new = key.shape[2]
        zeros = torch.zeros((key.shape[0], 1, n_new), dtype=torch.float32)
        fresh_life = zeros + 1e-7                           # kv_memory_store.py:38
        if self.key is None:
            self.key, self.shrinkage, self.selection = key, shrinkage, selection
            if self.count_usage:
                self.use_count, self.life_count = zeros, fresh_life
        else:
            self.key = torch.cat([self.key, key], dim=-1)
            if shrinkage is not None:
                self.shrinkage = torch.cat([self.shrinkage, shrinkage], dim=-1)
            if selection is not None:
                self.selection = torch.cat([self.selection, selection], dim=-1)
            if self.count_usage:
                self.use_count = torch.cat([self.use_count, zeros], dim=-1)
                self.life_count = torch.cat([self.life_count, fresh_life], dim=-1)

        if objects is not None:
            # working memory: ``value`` is one tensor indexed by (object id - 1)
            pending = [o - 1 for o in objects]
            for gi, members in enumerate(self.groups):
                for o in members:
                    pending.remove(o)
                self.values[gi] = torch.cat([self.values[gi], value[members]], dim=-1)
            if pending:
                self.values.append(value[pending])
                self.groups.append(list(pending))
                self.seen.extend(pending)
                if sorted(self.seen) != self.seen:          # kv_memory_store.py:79
                    raise AssertionError('Objects MUST be inserted in sorted order ')
        else:
            # long-term memory: ``value`` is already a per-group list
            for gi, gv in enumerate(value):
                if gv is None:
                    continue
                if gi < self.num_groups:
                    self.values[gi] = torch.cat([self.values[gi], gv], dim=-1)
                else:
                    self.values.append(gv)

    # -- usage (kv_memory_store.py:92-99,158-164) ----------------------------
    def record_usage(self, usage: Tensor) -> None:
        if not self.count_usage:
            return
        self.use_count = self.use_count + usage.reshape(self.use_count.shape)
        self.life_count = self.life_count + 1

    def normalized_usage(self) -> Tensor:
        if not self.count_usage:
            raise RuntimeError('I did not count usage!')
        return self.use_count / self.life_count

    # -- shrink (kv_memory_store.py:101-156) ----------------------------------
    def drop_range(self, start: int, end: int, min_size: int) -> None:
        """Keep [0,start) ++ [end,N) (``end`` <= 0 counts from the back; 0 = to the end)."""
        def cut(t: Tensor) -> Tensor:
            if end == 0:
                return t[:, :, :start]
            return torch.cat([t[:, :, :start], t[:, :, end:]], dim=-1)

        self.key = cut(self.key)
        if self.count_usage:
            self.use_count = cut(self.use_count)
            self.life_count = cut(self.life_count)
        if self.shrinkage is not None:
            self.shrinkage = cut(self.shrinkage)
        if self.selection is not None:
            self.selection = cut(self.selection)
        for gi in range(self.num_groups):
            if self.values[gi].shape[-1] >= min_size:
                self.values[gi] = cut(self.values[gi])

    def evict_least_used(self, max_size: int) -> None:
        usage = self.normalized_usage().flatten()
        lows, _ = torch.topk(usage, k=self.size - max_size, largest=False, sorted=True)
        keep = usage > lows[-1]                             # kv_memory_store.py:139-140
        if self.num_groups > 1:
            raise NotImplementedError('feature removal with multiple object groups is unsupported')
        self.key = self.key[:, :, keep]
        self.shrinkage = self.shrinkage[:, :, keep] if self.shrinkage is not None else None
        self.selection = self.selection[:, :, keep] if self.selection is not None else None
        for gi in range(self.num_groups):
            self.values[gi] = self.values[gi][:, :, keep]
        self.use_count = self.use_count[:, :, keep]
        self.life_count = self.life_count[:, :, keep]

    def window(self, start: int, end: int):
        """(key, shrinkage, selection, normalized usage) over [start, end) (kv_memory_store.py:166-181)."""
        sl = slice(start, None) if end == 0 else slice(start, end)
        usage = self.normalized_usage()[:, :, sl]
        pick = lambda t: None if t is None else t[:, :, sl]
        return pick(self.key), pick(self.shrinkage), pick(self.selection), usage


# ---------------------------------------------------------------------------
# memory_manager.py -- the hot path
# ---------------------------------------------------------------------------
@dataclass
class MatchResult:
    readout: Tensor                       # num_objects x CV x h x w
    work_usage: Optional[Tensor] = None   # N_work   (group-0 affinity row sums)
    long_usage: Optional[Tensor] = None   # N_long
    # per group: indices into that group's candidate axis, softmax weights, raw scores
    indices: List[Tensor] = field(default_factory=list)
    weights: List[Tensor] = field(default_factory=list)
    scores: List[Tensor] = field(default_factory=list)
    # per group: offset such that candidate j of the group is key (j + key_offset) of
    # the concatenated [long | work] timeline for the work part ... see ``group_key_index``
    group_spans: List[Tuple[int, int, int, int]] = field(default_factory=list)


def group_key_index(span: Tuple[int, int, int, int], j: Tensor) -> Tensor:
    """Map candidate positions of one group to positions on the [long | work] key axis.

    ``span`` = (long_size, long_len_g, work_size, work_len_g): the group's candidate axis is
    the last ``long_len_g`` long-term keys followed by the last ``work_len_g`` working keys
    (memory_manager.py:83,92-93,97).
    """
    n_long, len_l, n_work, len_w = span
    in_long = j < len_l
    return torch.where(in_long, j + (n_long - len_l), j - len_l + n_long + (n_work - len_w))


def match(work: Bank, long: Optional[Bank], q_key: Tensor, q_sel: Optional[Tensor], top_k: int,
          enable_long_term: bool, count_long_usage: bool, cv: int, sparse_only: bool = False) -> MatchResult:
    """``MemoryManager.match_memory`` (memory_manager.py:57-150) without the side effects.

    Usage row sums are returned instead of being added to the banks (the caller does
    that: ``Readout.match_memory``).  ``sparse_only`` skips the dense readout products and
    fills only indices / weights / scores (used for fp64 gap computation at large sizes).
    """
    h, w = q_key.shape[-2:]
    qk = q_key.flatten(start_dim=2)
    qe = q_sel.flatten(start_dim=2) if q_sel is not None else None
    n_groups = work.num_groups
    res = MatchResult(readout=torch.empty(0))

    use_long = enable_long_term and long is not None and long.engaged()
    if use_long:
        n_long = long.size
        keys = torch.cat([long.key, work.key], dim=-1)                  # memory_manager.py:73
        shr = torch.cat([long.shrinkage, work.shrinkage], dim=-1)       # memory_manager.py:74
        sim = anisotropic_l2(keys, shr, qk, qe)                         # memory_manager.py:76
        sim_work, sim_long = sim[:, n_long:], sim[:, :n_long]
        per_group_sim, per_group_val = [], []
        for gi in range(n_groups):
            len_w = work.group_len(gi)
            if gi < long.num_groups:
                len_l = long.group_len(gi)
                if gi == 0:
                    # memory_manager.py:83 -- long suffix of group 0 ++ ALL working keys
                    s = torch.cat([sim_long[:, -len_l:], sim_work], dim=1)
                    len_w = work.size
                else:
                    s = torch.cat([sim_long[:, -len_l:], sim_work[:, -len_w:]], dim=1)   # :92-93
                v = torch.cat([long.values[gi], work.values[gi]], dim=-1)               # :105
            else:
                len_l = 0
                if gi == 0:
                    s, len_w = sim_work, work.size
                else:
                    s = sim_work[:, -len_w:]                                             # :97
                v = work.values[gi]
            per_group_sim.append(s)
            per_group_val.append(v)
            res.group_spans.append((n_long, len_l, work.size, len_w))
    else:
        n_long = 0
        sim = anisotropic_l2(work.key, work.shrinkage, qk, qe)          # memory_manager.py:122
        per_group_sim, per_group_val = [], []
        for gi in range(n_groups):
            len_w = work.size if gi == 0 else work.group_len(gi)
            per_group_sim.append(sim if gi == 0 else sim[:, -len_w:])   # memory_manager.py:138
            per_group_val.append(work.values[gi])
            res.group_spans.append((0, 0, work.size, len_w))

    want_usage = use_long or enable_long_term                           # memory_manager.py:82,124
    outs = []
    for gi, (s, v) in enumerate(zip(per_group_sim, per_group_val)):
        idx, wts, vals = topk_select(s, top_k)
        res.indices.append(idx[0])
        res.weights.append(wts[0])
        res.scores.append(vals[0])
        aff = None
        if not sparse_only:
            aff = torch.zeros_like(s).scatter_(1, idx, wts)             # memory_util.py:51,54
        if gi == 0 and want_usage:
            if aff is not None:
                usage = aff.sum(dim=2).flatten()                        # memory_util.py:63
            else:
                usage = torch.zeros_like(s[:, :, 0]).scatter_add_(1, idx.flatten(1), wts.flatten(1)).flatten()
            if use_long:
                # group-0 candidate axis = long suffix ++ work; the reference slices at
                # long_size (memory_manager.py:113,118), which only lines up when group 0
                # covers every long-term key -- always true in the reference's lifecycle.
                len_l = res.group_spans[0][1]
                res.work_usage = usage[len_l:]
                if count_long_usage:
                    full = torch.zeros(n_long, dtype=usage.dtype)
                    full[n_long - len_l:] = usage[:len_l]
                    res.long_usage = full
            else:
                res.work_usage = usage
        if aff is not None:
            outs.append(torch.matmul(v, aff))                           # memory_manager.py:55,146
    if not sparse_only:
        stacked = torch.cat(outs, dim=0)
        res.readout = stacked.reshape(stacked.shape[0], cv, h, w)       # memory_manager.py:150
    return res


class Readout:
    """Restatement of ``MemoryManager`` (memory_manager.py:8-286) on top of ``Bank``."""

    def __init__(self, config: dict):
        self.hidden_dim = config['hidden_dim']
        self.top_k = config['top_k']
        self.enable_long_term = config['enable_long_term']
        self.enable_long_term_usage = config['enable_long_term_count_usage']
        if self.enable_long_term:
            self.max_mt_frames = config['max_mid_term_frames']
            self.min_mt_frames = config['min_mid_term_frames']
            self.num_prototypes = config['num_prototypes']
            self.max_long_elements = config['max_long_term_elements']
        self.CK = self.CV = None
        self.H = self.W = None
        self.hidden = None
        self.work_mem = Bank(count_usage=self.enable_long_term)
        self.long_mem = Bank(count_usage=self.enable_long_term_usage) if self.enable_long_term else None
        self._fresh = True

    # hot path ---------------------------------------------------------------
    def match_memory(self, query_key: Tensor, selection: Optional[Tensor]) -> Tensor:
        r = match(self.work_mem, self.long_mem, query_key, selection, self.top_k,
                  self.enable_long_term, self.enable_long_term_usage, self.CV)
        if r.work_usage is not None:
            self.work_mem.record_usage(r.work_usage)                    # memory_manager.py:114,129
        if r.long_usage is not None:
            self.long_mem.record_usage(r.long_usage)                    # memory_manager.py:119
        return r.readout

    # growth -----------------------------------------------------------------
    def add_memory(self, key, shrinkage, value, objects, selection=None) -> None:
        if self.H is None or self._fresh:
            self._fresh = False
            self.H, self.W = key.shape[-2:]
            self.HW = self.H * self.W
            if self.enable_long_term:
                self.min_work_elements = self.min_mt_frames * self.HW   # memory_manager.py:162
                self.max_work_elements = self.max_mt_frames * self.HW   # memory_manager.py:163
        key = key.flatten(start_dim=2)
        shrinkage = shrinkage.flatten(start_dim=2)
        value = value[0].flatten(start_dim=2)
        self.CK, self.CV = key.shape[1], value.shape[1]
        if selection is not None:
            selection = selection.flatten(start_dim=2)
        self.work_mem.append(key, value, shrinkage, selection, objects)
        if self.enable_long_term and self.work_mem.size >= self.max_work_elements:
            cap = self.max_long_elements - self.num_prototypes
            if self.long_mem.size >= cap:                               # memory_manager.py:187-188
                self.long_mem.evict_least_used(cap)
            self._compress()

    def _compress(self) -> None:
        """memory_manager.py:211-241: fold the middle of working memory into prototypes."""
        hw, total = self.HW, self.work_mem.size
        lo, hi = hw, -self.min_work_elements + hw
        cand_values = []
        for gv in self.work_mem.values:
            n_g = gv.shape[-1]
            if n_g == total or n_g > self.min_work_elements + hw:
                cand_values.append(gv[:, :, lo:hi])
            else:
                assert hw <= n_g < total
                cand_values.append(None)
        pk, pv, ps = self._consolidate(*self.work_mem.window(lo, hi), cand_values)
        self.work_mem.drop_range(lo, hi, min_size=self.min_work_elements + hw)
        self.long_mem.append(pk, pv, ps, selection=None, objects=None)

    def _consolidate(self, cand_key, cand_shrink, cand_sel, usage, cand_values):
        """memory_manager.py:245-286: top-usage prototypes + potentiation readout."""
        n = cand_key.shape[-1]
        _, top = torch.topk(usage, k=self.num_prototypes, dim=-1, sorted=True)
        proto = top.flatten()
        valid = [proto >= (n - gv.shape[2]) if gv is not None else None for gv in cand_values]
        proto_key = cand_key[:, :, proto]
        proto_sel = cand_sel[:, :, proto] if cand_sel is not None else None
        sim = anisotropic_l2(cand_key, cand_shrink, proto_key, proto_sel)
        affs = []
        for gi, gv in enumerate(cand_values):
            if gv is None:
                affs.append(None)
                continue
            a = dense_affinity(sim[:, -gv.shape[2]:, valid[gi]])
            affs.append(a if a.shape[-1] > 0 else None)
        proto_val = [torch.matmul(gv, affs[gi]) if affs[gi] is not None else None
                     for gi, gv in enumerate(cand_values)]
        proto_shrink = torch.matmul(cand_shrink, affs[0]) if cand_shrink is not None else None
        return proto_key, proto_val, proto_shrink

    # sensory memory -----------------------------------------------------------
    def create_hidden_state(self, n: int, sample_key: Tensor) -> None:
        h, w = sample_key.shape[-2:]
        if self.hidden is None:
            self.hidden = torch.zeros((1, n, self.hidden_dim, h, w))
        elif self.hidden.shape[1] != n:
            extra = torch.zeros((1, n - self.hidden.shape[1], self.hidden_dim, h, w))
            self.hidden = torch.cat([self.hidden, extra], dim=1)
        assert self.hidden.shape[1] == n

    def set_hidden(self, hidden):
        self.hidden = hidden

    def get_hidden(self):
        return self.hidden


# ---------------------------------------------------------------------------
# parity rule helpers (BASELINE.json north_star)
# ---------------------------------------------------------------------------
def key_projection(x: Tensor, key_w: Tensor, key_b: Tensor, d_w: Tensor, d_b: Tensor, e_w: Tensor, e_b: Tensor,
                   need_s: bool = True, need_e: bool = True):
    """KeyProjection.forward (tracker/model/modules.py:206-211): three 3x3 convolutions with padding 1 of
    x (B x in_dim x h x w); shrinkage = d_proj(x)^2 + 1, selection = sigmoid(e_proj(x)).  Works in the dtype of x
    (fp32 = the contract, fp64 for error measurements)."""
    conv = torch.nn.functional.conv2d
    key = conv(x, key_w.to(x.dtype), key_b.to(x.dtype), padding=1)
    shrinkage = conv(x, d_w.to(x.dtype), d_b.to(x.dtype), padding=1) ** 2 + 1 if need_s else None
    selection = torch.sigmoid(conv(x, e_w.to(x.dtype), e_b.to(x.dtype), padding=1)) if need_e else None
    return key, shrinkage, selection


def topk_gap(sim64: Tensor, k: int) -> Tensor:
    """Gap between the k-th and (k+1)-th largest fp64 similarity per query column."""
    kk = min(k + 1, sim64.shape[1])
    top = torch.topk(sim64, k=kk, dim=1).values
    if kk <= k:
        return torch.full_like(top[:, 0], float('inf'))
    return top[:, k - 1] - top[:, k]


def index_sets_equal(idx_a: Tensor, idx_b: Tensor) -> Tensor:
    """Per-query boolean: do two k x HW index tensors hold the same *set* per column?"""
    sa = torch.sort(idx_a.to(torch.int64), dim=0).values
    sb = torch.sort(idx_b.to(torch.int64), dim=0).values
    return (sa == sb).all(dim=0)


def rel_err(got: Tensor, want: Tensor) -> float:
    """Norm-wise relative error max|got-want| / max|want| (the 1e-2 tolerance is on this)."""
    denom = float(want.abs().max())
    if denom == 0.0:
        return float((got - want).abs().max())
    return float((got.double() - want.double()).abs().max()) / denom
