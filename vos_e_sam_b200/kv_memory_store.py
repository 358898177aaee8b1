"""Drop-in for the reference's ``KeyValueMemoryStore`` (tracker/inference/kv_memory_store.py), laid out
for the B200 kernels.

Same public surface (``add``, ``update_usage``, ``sieve_by_range``, ``remove_obsolete_features``,
``get_usage``, ``get_all_sliced``, ``get_v_size``, ``engaged``, ``size``, ``num_groups``, ``key`` /
``value`` / ``shrinkage`` / ``selection``), different storage:

* every tensor lives in a preallocated, geometrically grown device buffer; ``add`` writes only the
  appended range instead of re-concatenating the whole bank (kv_memory_store.py:49-56,70,88);
* next to the reference-layout fp32 tensors (CK x N keys, n_g x CV x N values -- still what ``key`` /
  ``value`` return, as views) the store keeps what the kernels read:
    - ``image``  : keys pre-multiplied by shrinkage / sqrt(CK), squared and split into bf16 hi/lo, tiled
                   64 keys at a time in the tcgen05 shared-memory operand layout (one bulk copy per tile)
    - ``shadow`` : per object group, values transposed to one memory element per row (bf16 or fp32), so
                   the sparse readout gathers contiguous rows.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import ops
from ._native import KEY_TILE

MIN_CAPACITY = 4096


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class _Grow:
    """A [lead..., capacity] fp32 buffer with a logical length on the last axis."""

    def __init__(self, lead: tuple, device, fill: float = 0.0, min_capacity: int = 0):
        self.lead, self.device, self.fill = lead, device, fill
        self.min_capacity = min_capacity            # first allocation is at least this large (KeyValueMemoryStore.reserve)
        self.buf = torch.empty(lead + (0,), dtype=torch.float32, device=device)
        self.n = 0

    @property
    def capacity(self) -> int:
        return self.buf.shape[-1]

    def reserve(self, total: int) -> bool:
        if total <= self.capacity:
            return False
        cap = _round_up(max(total, 2 * self.capacity, MIN_CAPACITY, self.min_capacity), KEY_TILE)
        new = torch.empty(self.lead + (cap,), dtype=torch.float32, device=self.device)
        if self.n:
            new[..., :self.n].copy_(self.buf[..., :self.n])
        self.buf = new
        return True

    def view(self) -> torch.Tensor:
        return self.buf[..., :self.n]

    def append(self, t: torch.Tensor) -> None:
        m = t.shape[-1]
        self.reserve(self.n + m)
        self.buf[..., self.n:self.n + m].copy_(t)
        self.n += m

    def append_const(self, m: int, value: float) -> None:
        self.reserve(self.n + m)
        self.buf[..., self.n:self.n + m].fill_(value)
        self.n += m

    def cut(self, start: int, stop: int) -> None:
        """Remove [start, stop) on the last axis."""
        tail = self.n - stop
        if tail > 0 and stop > start:
            src = self.buf[..., stop:self.n]
            if tail > stop - start:          # source and destination overlap: stage the tail
                src = src.clone()
            self.buf[..., start:start + tail].copy_(src)
        self.n -= (stop - start)

    def keep(self, index: torch.Tensor) -> None:
        kept = self.buf[..., :self.n].index_select(-1, index)
        self.n = kept.shape[-1]
        self.buf[..., :self.n].copy_(kept)


class _ValueGroup:
    """Values of one object group: fp32 n_g x CV x N (reference layout) + the row-per-element shadow."""

    def __init__(self, n_obj: int, cv: int, device, shadow_dtype, min_capacity: int = 0):
        self.n_obj, self.cv = n_obj, cv
        self.ref = _Grow((n_obj, cv), device, min_capacity=min_capacity)
        self.shadow_dtype = shadow_dtype
        self.shadow = torch.empty((0, n_obj * cv), dtype=shadow_dtype, device=device)

    @property
    def n(self) -> int:
        return self.ref.n

    @property
    def rows(self) -> int:
        return self.n_obj * self.cv

    def _sync_shadow(self, begin: int) -> None:
        """(Re)build shadow rows [begin, n) from the reference-layout buffer."""
        if self.shadow.shape[0] < self.ref.capacity:
            new = torch.empty((self.ref.capacity, self.rows), dtype=self.shadow_dtype, device=self.ref.device)
            if begin:
                new[:begin].copy_(self.shadow[:begin])
            self.shadow = new
        if self.n > begin:
            flat = self.ref.buf.view(self.rows, self.ref.capacity)
            ops.pack_values(flat, begin, self.n - begin, self.shadow, begin)

    def append(self, v: torch.Tensor) -> None:
        begin = self.n
        self.ref.append(v)
        self._sync_shadow(begin)

    def cut(self, start: int, stop: int) -> None:
        tail = self.n - stop
        self.ref.cut(start, stop)
        if 0 < tail <= stop - start and self.shadow.shape[0] >= self.ref.capacity:
            # the shadow rows of the tail move as they are (no overlap): no re-transposition of the fp32 values
            self.shadow[start:start + tail].copy_(self.shadow[stop:stop + tail])
        else:
            self._sync_shadow(start)

    def keep(self, index: torch.Tensor) -> None:
        self.ref.keep(index)
        self._sync_shadow(0)


class KeyValueMemoryStore:
    """Key / value bank shared by working and long-term memory (kv_memory_store.py:4-214)."""

    def __init__(self, count_usage: bool, value_dtype: torch.dtype = torch.bfloat16):
        self.count_usage = count_usage
        self.value_dtype = value_dtype
        self._k: Optional[_Grow] = None
        self._s: Optional[_Grow] = None
        self._e: Optional[_Grow] = None
        self._use: Optional[_Grow] = None
        self._life: Optional[_Grow] = None
        self._image: Optional[torch.Tensor] = None
        self._groups: List[_ValueGroup] = []
        self.obj_groups: List[List[int]] = []
        self.all_objects: List[int] = []
        self._min_capacity = 0
        self._version = 0     # bumped by everything that changes sizes or buffers (MemoryManager caches its descriptors on it)

    def reserve(self, n_elements: int) -> None:
        """Size every buffer that is allocated from now on for at least `n_elements` memory elements.  With the bank's
        bound known up front (working memory: max_mid_term_frames + 1 frames; long-term: max_long_term_elements) the
        buffers are allocated once and never move, so device pointers captured in a CUDA graph or a descriptor stay
        valid over add / sieve / eviction (the reference re-concatenates the bank on every add: kv_memory_store.py:49-56)."""
        self._min_capacity = max(self._min_capacity, int(n_elements))

    # ---- packed key image -----------------------------------------------------------------------
    def _sync_image(self, begin: int) -> None:
        """(Re)pack keys [begin, size) into the tensor-core image (only defined for CK == 64)."""
        k = self._k
        if k.lead[1] != 64:
            self._image = None
            return
        need = ops.key_image_bytes(64, k.capacity)
        if self._image is None or self._image.numel() < need:
            # capacity grew (geometrically, so rarely): start a fresh image and repack everything
            self._image = torch.empty(need, dtype=torch.uint8, device=k.device)
            begin = 0
        if k.n > begin:
            shr = self._s.buf.view(-1) if self._s is not None else None
            ops.pack_keys(k.buf[0], shr, begin, k.n, self._image, k.capacity)

    # ---- growth (kv_memory_store.py:36-90) ---------------------------------------------------------
    def add(self, key, value, shrinkage, selection, objects: Optional[List[int]]):
        ops._need(key, 'key')
        self._version += 1
        if self._append_fused(key, value, shrinkage, selection, objects):
            return
        device, m = key.device, key.shape[2]
        first = self._k is None
        if first:
            mc = self._min_capacity
            self._k = _Grow((key.shape[0], key.shape[1]), device, min_capacity=mc)
            self._s = _Grow((key.shape[0], 1), device, min_capacity=mc) if shrinkage is not None else None
            self._e = _Grow((key.shape[0], key.shape[1]), device, min_capacity=mc) if selection is not None else None
            if self.count_usage:
                self._use = _Grow((key.shape[0], 1), device, min_capacity=mc)
                self._life = _Grow((key.shape[0], 1), device, min_capacity=mc)
        begin = self._k.n
        self._k.append(key)
        if shrinkage is not None and self._s is not None:
            self._s.append(shrinkage)
        elif (shrinkage is None) != (self._s is None) and not first:
            raise RuntimeError('shrinkage must be given on every add or on none')
        if selection is not None and self._e is not None:
            self._e.append(selection)
        if self.count_usage:
            self._use.append_const(m, 0.0)       # kv_memory_store.py:37
            self._life.append_const(m, 1e-7)     # kv_memory_store.py:38
        self._sync_image(begin)

        if objects is not None:
            # working memory: `value` is one tensor indexed by (object id - 1); objects already in a
            # group extend that group, the rest open a new group (kv_memory_store.py:59-79)
            assert isinstance(value, torch.Tensor)
            remaining = [obj - 1 for obj in objects]
            for gi, group in enumerate(self.obj_groups):
                for obj in group:
                    remaining.remove(obj)  # ValueError on overlapping groups, like the reference
                self._groups[gi].append(value[group])
            if remaining:
                new_group = list(remaining)
                vg = _ValueGroup(len(new_group), value.shape[1], device, self.value_dtype, self._min_capacity)
                vg.append(value[new_group])
                self._groups.append(vg)
                self.obj_groups.append(new_group)
                self.all_objects.extend(new_group)
                assert sorted(self.all_objects) == self.all_objects, 'Objects MUST be inserted in sorted order '
        else:
            # long-term memory: `value` is a per-group list (kv_memory_store.py:80-90)
            assert isinstance(value, list)
            for gi, gv in enumerate(value):
                if gv is None:
                    continue
                if gi < self.num_groups:
                    self._groups[gi].append(gv)
                else:
                    vg = _ValueGroup(gv.shape[0], gv.shape[1], device, self.value_dtype, self._min_capacity)
                    vg.append(gv)
                    self._groups.append(vg)

    def _append_fused(self, key, value, shrinkage, selection, objects) -> bool:
        """The steady-state add -- every buffer exists and has room, the object groups are the ones the bank already
        has -- as ONE C call (vosmem_store_append: two launches) instead of a dozen copy / fill / pack launches.
        Returns False when the call does not fit (first add, new objects, growth): the general path above handles it."""
        import ctypes as C
        from . import _native as N
        k = self._k
        if k is None or key.shape[0] != 1 or k.lead[0] != 1:
            return False
        m, n = key.shape[2], k.n
        if (shrinkage is None) != (self._s is None) or (selection is None) != (self._e is None):
            return False
        bufs = [g for g in (k, self._s, self._e, self._use, self._life) if g is not None]
        if any(g.n != n or g.capacity < n + m for g in bufs):
            return False
        if k.lead[1] == 64 and (self._image is None or self._image.numel() < ops.key_image_bytes(64, k.capacity)):
            return False
        # the value blocks of the groups, without copies
        blocks = []
        if objects is not None:
            if not isinstance(value, torch.Tensor) or [o - 1 for o in objects] != self.all_objects:
                return False                      # new objects (a new group) or a different object list
            for group, vg in zip(self.obj_groups, self._groups):
                if group != list(range(group[0], group[0] + len(group))):
                    return False
                blocks.append((value[group[0]:group[0] + len(group)], vg))
        else:
            if not isinstance(value, list) or len(value) != self.num_groups or any(v is None for v in value):
                return False
            blocks = list(zip(value, self._groups))
        if len(blocks) > N.MAX_GROUPS:
            return False
        for v, vg in blocks:
            if v.shape[0] != vg.n_obj or v.shape[1] != vg.cv or v.shape[2] != m or not v.is_contiguous() or \
                    vg.ref.capacity < vg.n + m or vg.shadow.shape[0] < vg.ref.capacity or v.dtype != torch.float32:
                return False
        key2 = key[0] if key.stride(2) == 1 else key[0].contiguous()
        d = N.AppendDesc()
        d.ck, d.m, d.n = k.lead[1], m, n
        d.key, d.key_ld = key2.data_ptr(), key2.stride(0)
        keep = [key2]
        if shrinkage is not None:
            s2 = ops._need(shrinkage, 'shrinkage').reshape(-1).contiguous()
            keep.append(s2)
            d.shrinkage, d.bank_shrinkage = s2.data_ptr(), self._s.buf.data_ptr()
        if selection is not None:
            e2 = ops._need(selection, 'selection')[0]
            e2 = e2 if e2.stride(1) == 1 else e2.contiguous()
            keep.append(e2)
            d.selection, d.selection_ld, d.bank_selection = e2.data_ptr(), e2.stride(0), self._e.buf.data_ptr()
        d.bank_key, d.bank_ld = k.buf.data_ptr(), k.capacity
        if self._use is not None:
            d.bank_use, d.bank_life = self._use.buf.data_ptr(), self._life.buf.data_ptr()
        d.key_image = self._image.data_ptr() if k.lead[1] == 64 else None
        d.capacity = k.capacity
        d.value_dtype = ops.DTYPE_CODE[self.value_dtype]
        d.n_groups = len(blocks)
        for gi, (v, vg) in enumerate(blocks):
            g = d.group[gi]
            v2 = v.reshape(vg.rows, m)
            keep.append(v2)
            g.value, g.value_ld, g.rows = v2.data_ptr(), m, vg.rows
            g.ref, g.ref_ld = vg.ref.buf.data_ptr(), vg.ref.capacity
            g.shadow, g.shadow_ld, g.n = vg.shadow.data_ptr(), vg.shadow.stride(0), vg.n
        N.check(N.lib.vosmem_store_append(C.byref(d), ops._stream()), 'vosmem_store_append')
        for g in bufs:
            g.n = n + m
        for _, vg in blocks:
            vg.ref.n += m
        return True

    # ---- usage (kv_memory_store.py:92-99,158-164) -------------------------------------------------
    def update_usage(self, usage):
        """API parity only: the per-frame path accumulates usage inside the readout kernel."""
        if not self.count_usage:
            return
        self._use.view().add_(usage.view_as(self._use.view()))
        ops.age(self._life.buf, self._life.n)

    def age(self) -> None:
        """life_count += 1 (the second half of update_usage)."""
        if self.count_usage and self._life is not None:
            ops.age(self._life.buf, self._life.n)

    def get_usage(self):
        if not self.count_usage:
            raise RuntimeError('I did not count usage!')
        return self._use.view() / self._life.view()

    # ---- shrinking (kv_memory_store.py:101-156) ---------------------------------------------------
    def sieve_by_range(self, start: int, end: int, min_size: int):
        """Keep [0, start) ++ [end, N); `end` follows Python slicing on each tensor's own length
        (0 = to the end).  Values of groups shorter than `min_size` are left alone."""
        self._version += 1

        def bounds(n: int):
            lo = min(start, n)
            hi = n if end == 0 else slice(end, None).indices(n)[0]
            return lo, max(hi, lo)

        lo, hi = bounds(self._k.n)
        for g in (self._k, self._s, self._e, self._use, self._life):
            if g is not None:
                g.cut(lo, hi)
        self._sync_image(lo)
        for vg in self._groups:
            if vg.n >= min_size:
                vlo, vhi = bounds(vg.n)
                vg.cut(vlo, vhi)

    def remove_obsolete_features(self, max_size: int):
        self._version += 1
        usage = self.get_usage().flatten()
        values, _ = torch.topk(usage, k=(self.size - max_size), largest=False, sorted=True)
        survived = (usage > values[-1]).nonzero().flatten()
        if self.num_groups > 1:
            raise NotImplementedError('feature removal with multiple object groups is not supported: the survivor '
                                      'indices are per key but not every key has a value in every group')
        for g in (self._k, self._s, self._e, self._use, self._life):
            if g is not None:
                g.keep(survived)
        self._sync_image(0)
        for vg in self._groups:
            vg.keep(survived)

    def get_all_sliced(self, start: int, end: int):
        sl = slice(start, None) if end == 0 else slice(start, end)
        k = self.key[:, :, sl]
        sk = self.shrinkage[:, :, sl] if self._s is not None else None
        ek = self.selection[:, :, sl] if self._e is not None else None
        return k, sk, ek, self.get_usage()[:, :, sl]

    # ---- sizes / views ----------------------------------------------------------------------------
    def get_v_size(self, ni: int):
        return self._groups[ni].n

    def engaged(self):
        return self._k is not None

    @property
    def size(self):
        return 0 if self._k is None else self._k.n

    @property
    def num_groups(self):
        return len(self._groups)

    @property
    def key(self):
        return None if self._k is None else self._k.view()

    k = key

    @property
    def value(self):
        return [vg.ref.view() for vg in self._groups]

    v = value

    @property
    def shrinkage(self):
        return None if self._s is None else self._s.view()

    s = shrinkage

    @property
    def selection(self):
        return None if self._e is None else self._e.view()

    e = selection

    @property
    def use_count(self):
        return None if self._use is None else self._use.view()

    @property
    def life_count(self):
        return None if self._life is None else self._life.view()

    # ---- what the kernels read ----------------------------------------------------------------------
    def key_segment(self, begin: int, end: int) -> ops.KeySegment:
        return ops.KeySegment(key=self._k.buf[0], shrinkage=self._s.buf.view(-1) if self._s is not None else None,
                              image=self._image, begin=begin, end=end)

    def value_segment(self, gi: int, first: int, with_usage: bool, usage_offset: int = 0) -> ops.ValueSegment:
        vg = self._groups[gi]
        track = with_usage and self.count_usage
        use = self._use.buf.view(-1)[usage_offset:] if track else None
        life = self._life.buf.view(-1)[usage_offset:] if track else None     # aged by the same readout launch
        return ops.ValueSegment(shadow=vg.shadow, first=first, count=vg.n, use_count=use, life_count=life)

    def group_rows(self, gi: int) -> int:
        return self._groups[gi].rows
