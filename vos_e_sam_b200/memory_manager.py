"""Drop-in for the reference's ``MemoryManager`` (tracker/inference/memory_manager.py).

Same constructor (the XMem config dict), same methods and attributes that ``InferenceCore`` touches
(tracker/inference/inference_core.py:28,37,78,81,90,116,123,128,135): ``match_memory``, ``add_memory``,
``create_hidden_state``, ``set_hidden``, ``get_hidden``, ``update_config``, ``work_mem``, ``long_mem``,
``hidden``, ``CK``, ``CV``, ``H``, ``W``.

``match_memory`` -- the per-frame hot path -- never builds the N x HW similarity / affinity matrices
(memory_manager.py:76,82-99): per object group it issues ONE C call that runs
  pack query -> fused tcgen05 similarity + candidate selection -> split merge -> softmax + usage +
  sparse readout
over the [long-term | working] banks in place (no torch.cat of keys, similarities or values:
memory_manager.py:73-74,83,92-93,105).  Usage counters are updated inside the readout kernel.

Optional config keys (not in the reference's YAML): ``vosmem_value_dtype`` ('bf16' default | 'fp32') --
storage type of the gather-friendly value shadow; ``vosmem_path`` ('auto' | 'simt' | 'tcgen05');
``vosmem_shard`` = 'n' -- one tracker spread over the ranks of ``torch.distributed`` (one process per GPU): every
rank keeps the banks (the ranks run the same add_memory calls) but scans only its N-shard of [long-term | working]
keys per frame; candidates, readout and usage are exchanged as described in ``sharded.py`` (class
``ShardedMatch`` below); results equal the single-GPU manager's.
"""
from __future__ import annotations

import warnings

import torch

from . import _native as N
from . import ops
from .kv_memory_store import KeyValueMemoryStore
from .memory_util import do_softmax, get_similarity

_VALUE_DTYPES = {'bf16': torch.bfloat16, 'bfloat16': torch.bfloat16, 'fp32': torch.float32, 'float32': torch.float32}
_PATHS = {'auto': N.PATH_AUTO, 'simt': N.PATH_SIMT, 'tcgen05': N.PATH_TCGEN05}
MAX_WIDE_TOPK = 512       # csrc/dense.cu softmax_topk_any_kernel


class MemoryManager:
    """Manages working, long-term and sensory memory and the working -> long-term transition."""

    @staticmethod
    def _checked_top_k(config):
        """The fused per-frame path holds one survivor per lane of a warp: top_k <= 32 (XMem ships 30, config.yaml:8).
        Wider settings, up to the dense softmax kernel's 512, take `_match_wide` (the reference's own dense sequence on
        this library's dense kernels: correct, not fast).  Checked here rather than at the first match_memory."""
        top_k = int(config['top_k'])
        if not 1 <= top_k <= MAX_WIDE_TOPK:
            raise ValueError(f'vos_e_sam_b200.MemoryManager: top_k={top_k} outside [1, {MAX_WIDE_TOPK}] (fused readout up to '
                             f'{N.MAX_TOPK}, dense kernels up to {MAX_WIDE_TOPK}; see INTEGRATION.md "Limits")')
        return top_k

    def __init__(self, config):
        self.hidden_dim = config['hidden_dim']
        self.top_k = self._checked_top_k(config)

        self.enable_long_term = config['enable_long_term']
        self.enable_long_term_usage = config['enable_long_term_count_usage']
        if self.enable_long_term:
            self.max_mt_frames = config['max_mid_term_frames']
            self.min_mt_frames = config['min_mid_term_frames']
            self.num_prototypes = config['num_prototypes']
            self.max_long_elements = config['max_long_term_elements']

        self.value_dtype = _VALUE_DTYPES[str(config.get('vosmem_value_dtype', 'bf16')).lower()]
        self.path = _PATHS[str(config.get('vosmem_path', 'auto')).lower()]

        # dimensions are inferred from the first add_memory
        self.CK = self.CV = None
        self.H = self.W = None

        # sensory memory: one tensor for all objects, B x num_objects x CH x H x W
        self.hidden = None

        self.work_mem = KeyValueMemoryStore(count_usage=self.enable_long_term, value_dtype=self.value_dtype)
        if self.enable_long_term:
            self.long_mem = KeyValueMemoryStore(count_usage=self.enable_long_term_usage, value_dtype=self.value_dtype)

        self.reset_config = True
        self._scratch = None
        self._plan = None     # cached C descriptors of the last single-group match_memory (see _match_cached)
        # N-sharded over the ranks of torch.distributed (see the module docstring)
        self._sharded = ShardedMatch(self, config) if str(config.get('vosmem_shard', '')).lower() == 'n' else None

    def update_config(self, config):
        self.reset_config = True
        self.hidden_dim = config['hidden_dim']
        self.top_k = self._checked_top_k(config)

        assert self.enable_long_term == config['enable_long_term'], 'cannot update this'
        assert self.enable_long_term_usage == config['enable_long_term_count_usage'], 'cannot update this'

        self.enable_long_term_usage = config['enable_long_term_count_usage']
        if self.enable_long_term:
            self.max_mt_frames = config['max_mid_term_frames']
            self.min_mt_frames = config['min_mid_term_frames']
            self.num_prototypes = config['num_prototypes']
            self.max_long_elements = config['max_long_term_elements']

    def _readout(self, affinity, v):
        """Dense readout of one object group (memory_manager.py:53-55); used by consolidation."""
        n_obj, cv, n = v.shape
        if v.stride(2) == 1 and (n_obj == 1 or v.stride(0) == cv * v.stride(1)):
            # a slice of a bank along the key axis (compress_features): rows keep the bank's pitch, nothing is copied
            flat = v.as_strided((n_obj * cv, n), (v.stride(1), 1))
        else:
            flat = v.contiguous().view(n_obj * cv, n)
        return ops.readout_dense(flat, affinity[0]).view(n_obj, cv, -1)

    # ------------------------------------------------------------------------------------------------
    def _plan_match(self, query_key, selection, out=None):
        """The object groups of one match_memory call as independent problems (ops.MatchProblem) and the output they
        fill.  `out`: optional num_objects x C x h x w buffer with C >= CV (e.g. the decoder's concatenated
        readout + hidden input); the readout then lands in its channels [0, CV)."""
        work = self.work_mem
        num_groups = work.num_groups
        h, w = query_key.shape[-2:]
        hw = h * w
        if query_key.shape[0] != 1:
            raise RuntimeError('match_memory expects batch size 1 (inference_core.py:53)')
        qk = ops._need(query_key, 'query_key').flatten(start_dim=2)[0]
        qe = ops._need(selection, 'selection').flatten(start_dim=2)[0] if selection is not None else None

        use_long = self.enable_long_term and self.long_mem.engaged()
        long = self.long_mem if use_long else None
        n_work = work.size
        n_long = long.size if use_long else 0
        # usage is taken from the first group, which always sees every key (memory_manager.py:80-84,124-129)
        track_work = use_long or self.enable_long_term
        track_long = use_long and self.enable_long_term_usage

        rows_total = sum(work.group_rows(gi) for gi in range(num_groups))
        if out is None:
            flat = torch.empty((rows_total, hw), dtype=torch.float32, device=qk.device)
            result = flat.view(rows_total // self.CV, self.CV, h, w)
            grouped = None
        else:
            ops._need(out, 'out')
            if out.dim() == 5 and out.shape[0] == 1:
                out = out[0]
            if out.dim() != 4 or out.shape[0] * self.CV != rows_total or out.shape[1] < self.CV or \
                    tuple(out.shape[2:]) != (h, w) or not out.is_contiguous():
                raise RuntimeError(f'match_memory: `out` must be a contiguous num_objects x (>= CV) x h x w buffer, got '
                                   f'{tuple(out.shape)} for {rows_total // self.CV} objects, CV={self.CV}, {h}x{w}')
            result = out[:, :self.CV]
            grouped = out.view(out.shape[0], out.shape[1], hw)[:, :self.CV]      # objects x CV x HW, object pitch C * HW

        problems = []
        row0 = 0
        for gi in range(num_groups):
            segments, values = [], []
            first = 0
            if use_long and gi < long.num_groups:
                len_l = long.get_v_size(gi)                       # memory_manager.py:83,92
                segments.append(long.key_segment(n_long - len_l, n_long))
                values.append(long.value_segment(gi, 0, with_usage=(gi == 0 and track_long),
                                                 usage_offset=n_long - len_l))
                first = len_l
            len_w = n_work if gi == 0 else work.get_v_size(gi)    # memory_manager.py:83,93,97,138
            segments.append(work.key_segment(n_work - len_w, n_work))
            values.append(work.value_segment(gi, first, with_usage=(gi == 0 and track_work),
                                             usage_offset=n_work - len_w))
            rows = work.group_rows(gi)
            dst = flat[row0:row0 + rows] if grouped is None else grouped[row0 // self.CV:(row0 + rows) // self.CV]
            problems.append(ops.MatchProblem(qk, qe, segments, values, rows, dst))
            row0 += rows
        # life_count += 1 on every store whose usage is recorded (kv_memory_store.py:99) happens inside group 0's
        # readout launch (ValueSegment.life_count), so nothing is left to do after the kernels
        return problems, result

    def match_memory(self, query_key, selection):
        """query_key, selection: B x CK x H x W (B == 1)  ->  num_objects x CV x H x W  (memory_manager.py:57-150)."""
        return self.match_memory_into(query_key, selection, None)

    def match_memory_into(self, query_key, selection, out):
        """match_memory with a caller-provided destination (not in the reference): `out` is a contiguous
        [1 x] num_objects x (CV + CH) x H x W buffer -- the decoder's concatenated input (model/modules.py:232-233) --
        whose channels [0, CV) receive the readout directly; the returned tensor is that view (``None``: a fresh
        num_objects x CV x H x W tensor).  See `readout_with_hidden`."""
        if self.top_k > N.MAX_TOPK:
            return self._match_wide(query_key, selection, out)
        if self._sharded is not None:
            return self._sharded.match(query_key, selection, out)
        if out is None:
            fast = self._match_cached(query_key, selection)
            if fast is not None:
                return fast
        problems, out = self._plan_match(query_key, selection, out)
        hw = out.shape[-2] * out.shape[-1]
        if self._scratch is None or self._scratch[0].shape != (hw, self.top_k) or self._scratch[0].device != out.device:
            self._scratch = (torch.empty((hw, self.top_k), dtype=torch.float32, device=out.device),
                             torch.empty((hw, self.top_k), dtype=torch.int64, device=out.device))
        if len(problems) > 1 and self.CK == 64 and self.path != N.PATH_SIMT:
            # several object groups (objects that appeared in different frames): same query, different key suffixes --
            # one selection launch and one readout launch for all of them
            ops.match_batch(problems, self.top_k)
            return out
        for p in problems:
            ops.match(p.qk, p.qe, p.segments, p.values, p.rows, self.top_k, out=p.out, path=self.path,
                      scratch=self._scratch)
        return out

    def _match_wide(self, query_key, selection, out):
        """top_k in (32, 512]: the reference's dense sequence (memory_manager.py:57-150) -- similarity over
        [long-term | working] keys, per-group top-k softmax over the group's key suffixes, usage from group 0, dense
        readout -- on csrc/dense.cu's kernels.  Materialises N x HW fp32 like the reference does."""
        work = self.work_mem
        h, w = query_key.shape[-2:]
        if query_key.shape[0] != 1:
            raise RuntimeError('match_memory expects batch size 1 (inference_core.py:53)')
        qk = query_key.flatten(start_dim=2)
        qe = selection.flatten(start_dim=2) if selection is not None else None
        use_long = self.enable_long_term and self.long_mem.engaged()
        if use_long:
            long = self.long_mem
            n_long = long.size
            sim = get_similarity(torch.cat([long.key, work.key], -1), torch.cat([long.shrinkage, work.shrinkage], -1), qk, qe)
        else:
            n_long = 0
            sim = get_similarity(work.key, work.shrinkage, qk, qe)
        n_work = work.size
        readouts = []
        for gi, gv in enumerate(work.value):
            len_l = long.get_v_size(gi) if use_long and gi < long.num_groups else 0
            len_w = n_work if gi == 0 else work.get_v_size(gi)
            if len_l == n_long and len_w == n_work:
                group_sim = sim
            else:       # the group's candidates: a suffix of each bank (memory_manager.py:83,92-97)
                group_sim = torch.cat([sim[:, n_long - len_l:n_long], sim[:, n_long + n_work - len_w:]], 1)
            want_usage = gi == 0 and (use_long or self.enable_long_term)
            res = do_softmax(group_sim, top_k=self.top_k, inplace=group_sim is not sim or work.num_groups == 1,
                             return_usage=want_usage)
            if want_usage:
                aff, usage = res
                work.update_usage(usage[:, len_l:].flatten())
                if use_long and self.enable_long_term_usage:
                    usage_long = usage.new_zeros(n_long)
                    usage_long[n_long - len_l:] = usage[0, :len_l]
                    long.update_usage(usage_long)
            else:
                aff = res
            value = torch.cat([long.value[gi], gv], -1) if len_l else gv
            readouts.append(self._readout(aff, value))
        result = torch.cat(readouts, 0).view(-1, self.CV, h, w)
        if out is None:
            return result
        if out.dim() == 5 and out.shape[0] == 1:
            out = out[0]
        out[:, :self.CV].copy_(result)
        return out[:, :self.CV]

    def _match_cached(self, query_key, selection):
        """The common per-frame call -- one object group, fresh output -- with the C descriptors of the previous frame:
        between two add_memory calls (mem_every - 1 of mem_every frames) only the query and the output pointers
        change, so the ~60 descriptor fields, the segment objects and the workspace lookup are not rebuilt.  Stores
        bump `_version` whenever sizes or buffers change.  Returns None when the call does not fit (several groups)."""
        import ctypes as C
        work = self.work_mem
        if work.num_groups != 1 or query_key.shape[0] != 1:
            return None
        use_long = self.enable_long_term and self.long_mem.engaged()
        h, w = query_key.shape[-2:]
        device = query_key.device
        stream = torch.cuda.current_stream(device).cuda_stream
        key = (work._version, self.long_mem._version if use_long else -1, h, w, stream, selection is None, self.top_k,
               self.path, device)
        plan = self._plan
        if plan is None or plan[0] != key:
            problems, _ = self._plan_match(query_key, selection, None)
            if len(problems) != 1:
                return None
            p = problems[0]
            keep: list = []
            sd = ops._select_desc(p.qk, p.qe, p.segments, self.top_k, 0, self.path, keep)
            rd = ops._readout_desc(sd.hw, self.top_k, p.rows, p.values, p.out, None)
            plan = self._plan = (key, sd, rd, p.rows, None)   # (the descriptors point into the stores' own buffers)
        _, sd, rd, rows, _ = plan
        qk = ops._need(query_key, 'query_key').flatten(start_dim=2)[0].contiguous()
        qe = ops._need(selection, 'selection').flatten(start_dim=2)[0].contiguous() if selection is not None else None
        out = torch.empty((rows, h * w), dtype=torch.float32, device=device)
        sd.query_key, sd.query_selection = qk.data_ptr(), (qe.data_ptr() if qe is not None else None)
        rd.out, rd.out_ld = out.data_ptr(), h * w
        N.check(N.lib.vosmem_match(C.byref(sd), C.byref(rd), None, None, stream), 'vosmem_match')
        return out.view(rows // self.CV, self.CV, h, w)

    def readout_with_hidden(self, query_key, selection):
        """``torch.cat([match_memory(...).unsqueeze(0), get_hidden()], 2)`` -- what the decoder consumes
        (inference_core.py:78-81, model/modules.py:232-233) -- without the copy of the readout: the kernel writes
        channels [0, CV) of the 1 x num_objects x (CV + CH) x H x W result in place, only the hidden state is copied."""
        hidden = self.hidden
        n_obj, ch = hidden.shape[1], hidden.shape[2]
        h, w = query_key.shape[-2:]
        fused = torch.empty((1, n_obj, self.CV + ch, h, w), dtype=torch.float32, device=hidden.device)
        fused[:, :, self.CV:] = hidden
        self.match_memory_into(query_key, selection, fused)
        return fused

    # ------------------------------------------------------------------------------------------------
    def add_memory(self, key, shrinkage, value, objects, selection=None):
        """key 1 x CK x H x W, shrinkage 1 x 1 x H x W, value 1 x num_objects x CV x H x W (memory_manager.py:152-190)."""
        if self.H is None or self.reset_config:
            self.reset_config = False
            self.H, self.W = key.shape[-2:]
            self.HW = self.H * self.W
            if self.enable_long_term:
                # frames -> memory elements
                self.min_work_elements = self.min_mt_frames * self.HW
                self.max_work_elements = self.max_mt_frames * self.HW
                # both banks are bounded (memory_manager.py:184-190): allocate them once, so that their device
                # pointers never move (graphs / descriptors stay valid).  Unbounded test configurations keep growing
                # geometrically.
                if self.max_work_elements + self.HW <= 1 << 20:
                    self.work_mem.reserve(self.max_work_elements + self.HW)
                if self.max_long_elements + self.num_prototypes <= 1 << 20:
                    self.long_mem.reserve(self.max_long_elements + self.num_prototypes)

        if self._sharded is not None:
            self._sharded.sync_usage()     # consolidation / eviction below read the usage of ALL queries

        key = key.flatten(start_dim=2)
        shrinkage = shrinkage.flatten(start_dim=2)
        value = value[0].flatten(start_dim=2)

        self.CK = key.shape[1]
        self.CV = value.shape[1]

        if selection is not None:
            if not self.enable_long_term:
                warnings.warn('the selection factor is only needed in long-term mode', UserWarning)
            selection = selection.flatten(start_dim=2)

        self.work_mem.add(key, value, shrinkage, selection, objects)

        if self.enable_long_term and self.work_mem.size >= self.max_work_elements:
            # make room in long-term memory first, then fold the middle of working memory into prototypes
            room = self.max_long_elements - self.num_prototypes
            if self.long_mem.size >= room:
                self.long_mem.remove_obsolete_features(room)
            self.compress_features()

    def create_hidden_state(self, n, sample_key):
        """n is the TOTAL number of objects (memory_manager.py:192-203)."""
        h, w = sample_key.shape[-2:]
        if self.hidden is None:
            self.hidden = torch.zeros((1, n, self.hidden_dim, h, w), device=sample_key.device)
        elif self.hidden.shape[1] != n:
            grown = torch.zeros((1, n - self.hidden.shape[1], self.hidden_dim, h, w), device=sample_key.device)
            self.hidden = torch.cat([self.hidden, grown], 1)
        assert self.hidden.shape[1] == n

    def set_hidden(self, hidden):
        self.hidden = hidden

    def get_hidden(self):
        return self.hidden

    # ------------------------------------------------------------------------------------------------
    def compress_features(self):
        """Working -> long-term consolidation (memory_manager.py:211-241): everything except the first frame
        and the newest min_mt_frames - 1 frames becomes `num_prototypes` prototypes."""
        hw = self.HW
        lo, hi = hw, -self.min_work_elements + hw
        total = self.work_mem.size
        candidate_value = []
        for gv in self.work_mem.value:
            n_g = gv.shape[-1]
            if n_g == total or n_g > self.min_work_elements + hw:
                candidate_value.append(gv[:, :, lo:hi])
            else:
                # a group that entered late and is still too short to be consolidated
                assert hw <= n_g < total
                candidate_value.append(None)

        prototype_key, prototype_value, prototype_shrinkage = self.consolidation(
            *self.work_mem.get_all_sliced(lo, hi), candidate_value)

        self.work_mem.sieve_by_range(lo, hi, min_size=self.min_work_elements + hw)
        self.long_mem.add(prototype_key, prototype_value, prototype_shrinkage, selection=None, objects=None)

    def consolidation(self, candidate_key, candidate_shrinkage, candidate_selection, usage, candidate_value):
        """Pick the most-used candidates as prototypes and aggregate values into them by a dense softmax
        readout ("potentiation", memory_manager.py:245-286)."""
        n = candidate_key.shape[-1]
        _, top = torch.topk(usage, k=self.num_prototypes, dim=-1, sorted=True)
        proto = top.flatten()

        # a prototype is only valid for a group whose (suffix) extent contains it
        validity = [proto >= (n - gv.shape[2]) if gv is not None else None for gv in candidate_value]

        prototype_key = candidate_key[:, :, proto]
        prototype_selection = candidate_selection[:, :, proto] if candidate_selection is not None else None

        similarity = get_similarity(candidate_key, candidate_shrinkage, prototype_key, prototype_selection)

        affinity = []
        for gi, gv in enumerate(candidate_value):
            if gv is None:
                affinity.append(None)
                continue
            sub = similarity[:, -gv.shape[2]:, validity[gi]]
            affinity.append(do_softmax(sub) if sub.shape[-1] > 0 else None)

        prototype_value = [self._readout(affinity[gi], gv) if affinity[gi] is not None else None
                           for gi, gv in enumerate(candidate_value)]
        prototype_shrinkage = (self._readout(affinity[0], candidate_shrinkage)
                               if candidate_shrinkage is not None else None)
        return prototype_key, prototype_value, prototype_shrinkage


class ShardedMatch:
    """``match_memory`` of ONE tracker spread over the ranks of ``torch.distributed`` (config ``vosmem_shard='n'``;
    SURVEY.md section 8e; the reference is single-GPU, tools/runner.py:32).

    Every rank holds the banks (the ranks run the same ``add_memory`` / consolidation calls on the same inputs, so
    the replicas stay identical) but per frame scans only ITS 64-key aligned blocks of the long-term and the working
    keys -- the two candidate segments of memory_manager.py:73-74, cut down per rank (``sharded.plan_shard_ranges``;
    groups that entered late see suffixes, memory_manager.py:88-98).  ``ShardedLongTermReadout`` moves the candidate
    lists to the ranks that own the queries, which merge them and read out their query slice from the (replicated)
    values; the slices are gathered so that every rank returns the full ``num_objects x CV x h x w`` readout.

    Usage (memory_manager.py:113-119): a rank's readout kernel adds the weights of ITS queries into rank-local
    delta buffers; ``sync_usage`` all-reduces the deltas into ``use_count`` before anything reads the usage
    (``add_memory`` -> ``remove_obsolete_features`` / ``compress_features``), which is every mem_every-th frame and
    not per frame.  ``life_count`` is aged on every rank by the same kernel."""

    def __init__(self, manager: 'MemoryManager', config):
        import torch.distributed as dist
        from .sharded import ShardedLongTermReadout
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("vosmem_shard='n' needs an initialised torch.distributed process group (one rank per GPU)")
        self.m = manager
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = torch.device('cuda', torch.cuda.current_device())
        cfg = dict(config)
        cfg['vosmem_shard'] = 'n'
        self.engine = ShardedLongTermReadout(cfg, self.rank, self.world, self.device)
        self.engine.backend.value_dtype = manager.value_dtype
        self._delta = {}      # id(store) -> flat fp32 usage delta of this rank since the last sync

    def _delta_of(self, store):
        cap = store._use.capacity
        d = self._delta.get(id(store))
        if d is None or d[0].numel() < cap:
            if d is not None:
                self._sync_store(store, d[0])          # (capacity grew: fold what was accumulated first)
            d = (torch.zeros(cap, dtype=torch.float32, device=store._use.buf.device), store)
            self._delta[id(store)] = d
        return d[0]

    def _sync_store(self, store, delta):
        import torch.distributed as dist
        n = store._use.n
        if n:
            part = delta[:n]
            if self.world > 1:
                dist.all_reduce(part)
            store._use.view().view(-1).add_(part)
        delta.zero_()

    def sync_usage(self):
        """use_count += sum over ranks of the usage their query slices produced since the last call."""
        for delta, store in list(self._delta.values()):
            self._sync_store(store, delta)

    def match(self, query_key, selection, out=None):
        from .sharded import ShardProblem, plan_shard_ranges
        m = self.m
        work = m.work_mem
        h, w = query_key.shape[-2:]
        hw = h * w
        if query_key.shape[0] != 1:
            raise RuntimeError('match_memory expects batch size 1 (inference_core.py:53)')
        ops._need(query_key, 'query_key')
        use_long = m.enable_long_term and m.long_mem.engaged()
        long = m.long_mem if use_long else None
        n_work, n_long = work.size, (long.size if use_long else 0)
        track_work = use_long or m.enable_long_term
        track_long = use_long and m.enable_long_term_usage
        rows_total = sum(work.group_rows(gi) for gi in range(work.num_groups))
        if out is None:
            result = torch.empty((rows_total // m.CV, m.CV, h, w), dtype=torch.float32, device=query_key.device)
            dst = result.view(rows_total, hw)
        else:
            if out.dim() == 5 and out.shape[0] == 1:
                out = out[0]
            result = out[:, :m.CV]
            dst = None
        row0 = 0
        for gi in range(work.num_groups):
            has_long = use_long and gi < long.num_groups
            len_l = long.get_v_size(gi) if has_long else 0
            len_w = n_work if gi == 0 else work.get_v_size(gi)
            ranges = [(n_long - len_l, n_long), (n_work - len_w, n_work)] if has_long else [(n_work - len_w, n_work)]
            sizes = [n_long, n_work] if has_long else [n_work]
            mine, base0, seg0_len, base1, share_ok = plan_shard_ranges(ranges, sizes, self.world, self.rank)
            banks = [long, work] if has_long else [work]
            segments = [b.key_segment(lo, hi) for b, (lo, hi) in zip(banks, mine)]
            values, first = [], 0
            for b, (begin, end), tracked in zip(banks, ranges, [track_long, track_work] if has_long else [track_work]):
                vg = b._groups[gi]
                track = gi == 0 and tracked and b.count_usage
                values.append(ops.ValueSegment(
                    shadow=vg.shadow, first=first, count=vg.n,
                    use_count=self._delta_of(b)[begin:] if track else None,
                    life_count=b._life.buf.view(-1)[begin:] if track else None))
                first += end - begin
            rows = work.group_rows(gi)
            prob = ShardProblem(segments, base0, seg0_len, base1, values, rows, first, share_ok)
            full = self.engine.match(query_key, selection, gather=True, problem=prob)     # rows x HW
            q_lo, q_hi = self.engine.query_range(hw)
            if q_hi <= q_lo and gi == 0:        # this rank owns no query: its readout kernel (which ages) did not run
                for b, tracked in zip(banks, [track_long, track_work] if has_long else [track_work]):
                    if tracked:
                        b.age()
            if dst is not None:
                dst[row0:row0 + rows].copy_(full)
            else:
                result[row0 // m.CV:(row0 + rows) // m.CV].copy_(full.view(rows // m.CV, m.CV, h, w))
            row0 += rows
        return result


def match_memory_batch(managers, query_keys, selections):
    """``[m.match_memory(k, e) for m, k, e in zip(managers, query_keys, selections)]`` for INDEPENDENT managers (one
    per sequence: tools/runner.py:61-63), with the selection and readout kernels of all of them launched together
    (``vosmem_match_batch``: blockIdx.z = problem).  Managers must share CK == 64, the query size and top_k; anything
    else falls back to one call per manager."""
    managers = list(managers)
    if any(m.top_k > N.MAX_TOPK for m in managers):       # dense path, one manager at a time
        return [m.match_memory(k, e) for m, k, e in zip(managers, query_keys, selections)]
    plans = [m._plan_match(k, e) for m, k, e in zip(managers, query_keys, selections)]
    top_k = managers[0].top_k
    hw = plans[0][1].shape[-2] * plans[0][1].shape[-1]
    same = all(m.top_k == top_k and m.CK == 64 and m.path != N.PATH_SIMT for m in managers) and \
        all(o.shape[-2] * o.shape[-1] == hw for _, o in plans)
    if same and len(set(map(id, managers))) == len(managers):
        ops.match_batch([p for probs, _ in plans for p in probs], top_k)
    else:
        for m, (probs, out) in zip(managers, plans):
            for p in probs:
                ops.match(p.qk, p.qe, p.segments, p.values, p.rows, m.top_k, out=p.out, path=m.path)
    return [out for _, out in plans]
