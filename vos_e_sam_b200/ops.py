"""Tensor-level wrappers over the C ABI: device pointers in, device tensors out.

PyTorch is used for device memory and the current stream only; every arithmetic step is a kernel of
libvosmem.so.  All functions require CUDA fp32 tensors and raise otherwise (no CPU path).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _native as N
from ._native import check

Tensor = torch.Tensor
DTYPE_CODE = {torch.float32: N.F32, torch.bfloat16: N.BF16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need(t: Tensor, name: str, dtype=torch.float32) -> Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f'vos_e_sam_b200: `{name}` must be a CUDA tensor (this package has no CPU path)')
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f'vos_e_sam_b200: `{name}` must be {dtype}, got {t.dtype}')
    return t


def _p(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _rows_2d(t: Tensor, name: str) -> Tuple[Tensor, int]:
    """(tensor, row pitch) of a 2-D tensor whose last dim is contiguous; copies only if it is not."""
    if t.dim() != 2:
        raise RuntimeError(f'`{name}` must be 2-D, got {tuple(t.shape)}')
    if t.shape[1] > 1 and t.stride(1) != 1:
        t = t.contiguous()
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
    return t, ld


# ------------------------------------------------------------------------------------------------
@dataclass
class KeySegment:
    """Candidate keys [begin, end) of one bank.  `key` is CK x (>= end) with row pitch key.stride(0)."""
    key: Optional[Tensor]
    shrinkage: Optional[Tensor]          # flat, >= end elements, or None
    image: Optional[Tensor]              # packed operand image (uint8) or None
    begin: int
    end: int


@dataclass
class ValueSegment:
    """Values of candidates [first, first+count): shadow is (>= count) x rows, one memory element per row."""
    shadow: Tensor
    first: int
    count: int
    use_count: Optional[Tensor] = None   # flat fp32, element i <-> shadow row i
    life_count: Optional[Tensor] = None  # flat fp32; [0, count) += 1 inside the readout launch when given


_workspaces: Dict[Tuple[int, int, int, int, int], Tensor] = {}


def workspace_for(device: torch.device, ck: int, hw: int, slot: int = 0) -> Tensor:
    """Cached scratch per (device, CUDA stream, shape); `slot` gives the problems of one batch distinct workspaces.

    A workspace carries state between the selection kernel and its consumer (candidate lists, published thresholds,
    the device-side launch epoch), so calls that share one must be stream-ordered: the cache is keyed by the current
    stream, which gives trackers that run on different streams of one device their own scratch."""
    index = device.index if device.index is not None else torch.cuda.current_device()
    key = (index, torch.cuda.current_stream(device).cuda_stream, ck, hw, slot)
    ws = _workspaces.get(key)
    if ws is None:
        if torch.cuda.is_current_stream_capturing():
            # the one-time initialisation must not become a graph node: it would reset the launch epoch on every replay
            raise RuntimeError('vos_e_sam_b200: run the call once on this stream before capturing it into a CUDA graph '
                               '(its workspace is allocated and initialised on first use)')
        nbytes = N.lib.vosmem_workspace_bytes(ck, hw, 0)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        check(N.lib.vosmem_workspace_init(ws.data_ptr(), nbytes, _stream()), 'vosmem_workspace_init')
        _workspaces[key] = ws
    return ws


def workspace_status(ws: Tensor) -> None:
    """Raise if a tcgen05 launch on `ws` flagged a device-side error (synchronises the current stream)."""
    check(N.lib.vosmem_workspace_status(ws.data_ptr(), _stream()), 'vosmem_workspace_status')


def key_image_bytes(ck: int, capacity: int) -> int:
    return int(N.lib.vosmem_key_image_bytes(ck, capacity))


def pack_keys(key: Tensor, shrinkage: Optional[Tensor], begin: int, end: int, image: Tensor, capacity: int) -> None:
    """Pack keys [begin, end) of `key` (CK x >=end, last dim contiguous) into `image`."""
    _need(key, 'key')
    _need(image, 'image', torch.uint8)
    assert key.dim() == 2 and (key.shape[1] <= 1 or key.stride(1) == 1)
    if shrinkage is not None:
        _need(shrinkage, 'shrinkage')
        assert shrinkage.is_contiguous()
    check(N.lib.vosmem_pack_keys(key.data_ptr(), key.stride(0), _p(shrinkage), key.shape[0], begin, end,
                                 image.data_ptr(), capacity, _stream()), 'vosmem_pack_keys')


def pack_values(value: Tensor, src_begin: int, n: int, shadow: Tensor, dst_begin: int) -> None:
    """shadow[dst_begin + i, r] = value[r, src_begin + i]; value is rows x (>= src_begin+n), last dim contiguous."""
    _need(value, 'value')
    _need(shadow, 'shadow', None)
    assert value.dim() == 2 and (value.shape[1] <= 1 or value.stride(1) == 1)
    assert shadow.dim() == 2 and shadow.stride(1) == 1 and shadow.shape[1] >= value.shape[0]
    check(N.lib.vosmem_pack_values(value.data_ptr(), value.stride(0), value.shape[0], src_begin, n, shadow.data_ptr(),
                                   shadow.stride(0), dst_begin, DTYPE_CODE[shadow.dtype], _stream()),
          'vosmem_pack_values')


def age(life_count: Tensor, n: int) -> None:
    _need(life_count, 'life_count')
    check(N.lib.vosmem_age(life_count.data_ptr(), n, _stream()), 'vosmem_age')


# ------------------------------------------------------------------------------------------------
def _select_desc(qk: Tensor, qe: Optional[Tensor], segments: Sequence[KeySegment], top_k: int, index_base: int,
                 path: int, keep: list, slot: int = 0, d: Optional[N.SelectDesc] = None,
                 workspace: Optional[Tensor] = None) -> N.SelectDesc:
    _need(qk, 'query_key')
    ck, hw = qk.shape
    qk = qk.contiguous()
    keep.append(qk)
    if qe is not None:
        qe = _need(qe, 'query_selection').contiguous()
        keep.append(qe)
    d = N.SelectDesc() if d is None else d
    d.ck, d.hw, d.top_k = ck, hw, top_k
    d.query_key, d.query_selection = qk.data_ptr(), _p(qe)
    d.n_segments = len(segments)
    for i, s in enumerate(segments):
        g = d.seg[i]
        if s.key is not None:
            _need(s.key, 'segment.key')
            assert s.key.dim() == 2 and s.key.shape[0] == ck and (s.key.shape[1] <= 1 or s.key.stride(1) == 1)
            g.key, g.key_ld = s.key.data_ptr(), s.key.stride(0)
        if s.shrinkage is not None:
            _need(s.shrinkage, 'segment.shrinkage')
            g.shrinkage = s.shrinkage.data_ptr()
        if s.image is not None:
            g.key_image = s.image.data_ptr()
        g.begin, g.end = s.begin, s.end
    d.index_base = index_base
    d.path = path
    ws = workspace if workspace is not None else workspace_for(qk.device, ck, hw, slot)   # caller-owned: already initialised
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    return d


def select_topk(qk: Tensor, qe: Optional[Tensor], segments: Sequence[KeySegment], top_k: int, index_base: int = 0,
                path: int = N.PATH_AUTO, out: Optional[Tuple[Tensor, Tensor]] = None) -> Tuple[Tensor, Tensor]:
    """Fused similarity + per-query top-k.  Returns (score, index), both HW x top_k, best first (written into `out`
    when given: contiguous fp32 / int64 tensors of that shape, e.g. views of a peer-mapped exchange buffer)."""
    keep: list = []
    d = _select_desc(qk, qe, segments, top_k, index_base, path, keep)
    if out is not None:
        score, index = _need(out[0], 'out score'), _need(out[1], 'out index', torch.int64)
        assert score.shape == (d.hw, top_k) and index.shape == (d.hw, top_k) and score.is_contiguous() and index.is_contiguous()
    else:
        score = torch.empty((d.hw, top_k), dtype=torch.float32, device=qk.device)
        index = torch.empty((d.hw, top_k), dtype=torch.int64, device=qk.device)
    check(N.lib.vosmem_select_topk(C.byref(d), score.data_ptr(), index.data_ptr(), _stream()), 'vosmem_select_topk')
    return score, index


def merge_topk(scores: Tensor, indices: Tensor) -> Tuple[Tensor, Tensor]:
    """scores / indices: n_lists x HW x top_k (e.g. all-gathered per-rank candidates) -> HW x top_k."""
    _need(scores, 'scores')
    _need(indices, 'indices', torch.int64)
    scores, indices = scores.contiguous(), indices.contiguous()
    n_lists, hw, k = scores.shape
    out_s = torch.empty((hw, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((hw, k), dtype=torch.int64, device=scores.device)
    check(N.lib.vosmem_merge_topk(scores.data_ptr(), indices.data_ptr(), n_lists, hw, k, out_s.data_ptr(),
                                  out_i.data_ptr(), _stream()), 'vosmem_merge_topk')
    return out_s, out_i


def merge_topk_ptrs(score_ptrs: Sequence[int], index_ptrs: Sequence[int], hw: int, top_k: int,
                    device: torch.device) -> Tuple[Tensor, Tensor]:
    """merge_topk over lists given as raw device addresses (one HW x top_k fp32 / int64 pair per list), e.g. the
    peer-mapped exchange buffers of the other ranks: the kernel reads them over NVLink itself."""
    n = len(score_ptrs)
    sp = (C.c_void_p * n)(*score_ptrs)
    ip = (C.c_void_p * n)(*index_ptrs)
    out_s = torch.empty((hw, top_k), dtype=torch.float32, device=device)
    out_i = torch.empty((hw, top_k), dtype=torch.int64, device=device)
    check(N.lib.vosmem_merge_topk_ptrs(sp, ip, n, hw, top_k, out_s.data_ptr(), out_i.data_ptr(), _stream()),
          'vosmem_merge_topk_ptrs')
    return out_s, out_i


def _readout_desc(hw: int, top_k: int, rows: int, values: Sequence[ValueSegment], out: Tensor,
                  out_weight: Optional[Tensor], d: Optional[N.ReadoutDesc] = None) -> N.ReadoutDesc:
    d = N.ReadoutDesc() if d is None else d
    d.hw, d.top_k, d.rows = hw, top_k, rows
    d.value_dtype = DTYPE_CODE[values[0].shadow.dtype]
    d.n_segments = len(values)
    for i, v in enumerate(values):
        _need(v.shadow, 'shadow', None)
        assert v.shadow.dtype == values[0].shadow.dtype and v.shadow.stride(1) == 1 and v.shadow.shape[1] >= rows
        g = d.seg[i]
        g.shadow, g.shadow_ld = v.shadow.data_ptr(), v.shadow.stride(0)
        g.first, g.count = v.first, v.count
        g.use_count = _p(v.use_count)
        g.life_count = _p(v.life_count)
    if out.dim() == 3:     # grouped rows: objects x CV x HW view of a larger (objects x (CV + CH) x HW) buffer
        assert out.shape[0] * out.shape[1] == rows and out.shape[2] == hw and out.stride(2) == 1
        d.out, d.out_ld = out.data_ptr(), out.stride(1)
        d.out_group_rows, d.out_group_stride = out.shape[1], out.stride(0)
    else:
        assert out.dim() == 2 and out.shape == (rows, hw) and out.stride(1) == 1
        d.out, d.out_ld = out.data_ptr(), out.stride(0)
        d.out_group_rows, d.out_group_stride = 0, 0
    d.out_weight = _p(out_weight)
    return d


def softmax_readout(score: Tensor, index: Tensor, values: Sequence[ValueSegment], rows: int,
                    out: Optional[Tensor] = None, want_weight: bool = False):
    """Softmax over the survivors + usage scatter-add + sparse readout.  Returns out (rows x HW) [, weights]."""
    _need(score, 'score')
    _need(index, 'index', torch.int64)
    hw, k = score.shape
    if out is None:
        out = torch.empty((rows, hw), dtype=torch.float32, device=score.device)
    weight = torch.empty((hw, k), dtype=torch.float32, device=score.device) if want_weight else None
    d = _readout_desc(hw, k, rows, values, out, weight)
    check(N.lib.vosmem_softmax_readout(C.byref(d), score.data_ptr(), index.data_ptr(), _stream()),
          'vosmem_softmax_readout')
    return (out, weight) if want_weight else out


def match(qk: Tensor, qe: Optional[Tensor], segments: Sequence[KeySegment], values: Sequence[ValueSegment], rows: int,
          top_k: int, out: Optional[Tensor] = None, path: int = N.PATH_AUTO,
          scratch: Optional[Tuple[Tensor, Tensor]] = None) -> Tensor:
    """One object group of MemoryManager.match_memory in a single C call."""
    keep: list = []
    sd = _select_desc(qk, qe, segments, top_k, 0, path, keep)
    hw = sd.hw
    if out is None:
        out = torch.empty((rows, hw), dtype=torch.float32, device=qk.device)
    if scratch is None:
        scratch = (torch.empty((hw, top_k), dtype=torch.float32, device=qk.device),
                   torch.empty((hw, top_k), dtype=torch.int64, device=qk.device))
    rd = _readout_desc(hw, top_k, rows, values, out, None)
    check(N.lib.vosmem_match(C.byref(sd), C.byref(rd), scratch[0].data_ptr(), scratch[1].data_ptr(), _stream()),
          'vosmem_match')
    return out


@dataclass
class MatchProblem:
    """One entry of match_batch: what ops.match takes, for one sequence / object group."""
    qk: Tensor
    qe: Optional[Tensor]
    segments: Sequence[KeySegment]
    values: Sequence[ValueSegment]
    rows: int
    out: Optional[Tensor] = None


def match_batch(problems: Sequence[MatchProblem], top_k: int) -> List[Tensor]:
    """Independent match problems (same CK == 64, HW and top_k; e.g. one frame of each of several sequences) in one
    selection launch and one readout launch per chunk of N.MAX_BATCH.  Returns the outputs (rows_i x HW)."""
    outs: List[Tensor] = []
    for c0 in range(0, len(problems), N.MAX_BATCH):
        chunk = problems[c0:c0 + N.MAX_BATCH]
        n = len(chunk)
        keep: list = []
        sds = (N.SelectDesc * n)()
        rds = (N.ReadoutDesc * n)()
        for i, p in enumerate(chunk):
            _select_desc(p.qk, p.qe, p.segments, top_k, 0, N.PATH_TCGEN05, keep, slot=i, d=sds[i])
            out = p.out if p.out is not None else torch.empty((p.rows, sds[i].hw), dtype=torch.float32, device=p.qk.device)
            _readout_desc(sds[i].hw, top_k, p.rows, p.values, out, None, d=rds[i])
            outs.append(out)
        check(N.lib.vosmem_match_batch(sds, rds, n, _stream()), 'vosmem_match_batch')
    return outs


# ------------------------------------------------------------------------------------------------
def similarity_dense(key: Tensor, shrinkage: Optional[Tensor], qk: Tensor, qe: Optional[Tensor]) -> Tensor:
    """key CK x N, shrinkage N or None, qk/qe CK x HW -> N x HW (memory_util.get_similarity for one batch item)."""
    key, key_ld = _rows_2d(_need(key, 'mk'), 'mk')
    qk = _need(qk, 'qk').contiguous()
    qe = _need(qe, 'qe').contiguous() if qe is not None else None
    if shrinkage is not None:
        shrinkage = _need(shrinkage, 'ms').contiguous()
    ck, n = key.shape
    hw = qk.shape[1]
    out = torch.empty((n, hw), dtype=torch.float32, device=key.device)
    check(N.lib.vosmem_similarity_dense(key.data_ptr(), key_ld, _p(shrinkage), qk.data_ptr(), _p(qe), ck, n, hw,
                                        out.data_ptr(), _stream()), 'vosmem_similarity_dense')
    return out


def softmax_dense(similarity: Tensor, top_k: Optional[int], inplace: bool, want_usage: bool):
    """similarity N x HW -> affinity N x HW [, usage N]."""
    given = similarity
    similarity, sim_ld = _rows_2d(_need(similarity, 'similarity'), 'similarity')
    if inplace and similarity.data_ptr() != given.data_ptr():
        raise RuntimeError('softmax_dense(inplace=True) needs a similarity whose last dimension is contiguous '
                           '(the kernel would otherwise update a copy)')
    n, hw = similarity.shape
    aff = similarity if inplace else torch.empty((n, hw), dtype=torch.float32, device=similarity.device)
    aff_ld = sim_ld if inplace else hw
    usage = torch.empty(n, dtype=torch.float32, device=similarity.device) if want_usage else None
    scratch = None
    if top_k is None or top_k <= 0:     # dense softmax: statistics per (row chunk, column) let it run parallel over rows
        scratch = torch.empty(int(N.lib.vosmem_softmax_dense_scratch_bytes(n, hw)), dtype=torch.uint8, device=similarity.device)
    check(N.lib.vosmem_softmax_dense_ws(similarity.data_ptr(), sim_ld, n, hw, top_k if top_k is not None else 0,
                                        aff.data_ptr(), aff_ld, _p(usage), _p(scratch),
                                        scratch.numel() if scratch is not None else 0, _stream()), 'vosmem_softmax_dense')
    return aff, usage


def readout_dense(value: Tensor, affinity: Tensor) -> Tensor:
    """value rows x N, affinity N x HW -> rows x HW."""
    value, v_ld = _rows_2d(_need(value, 'value'), 'value')
    affinity, a_ld = _rows_2d(_need(affinity, 'affinity'), 'affinity')
    rows, n = value.shape
    if affinity.shape[0] != n:
        raise RuntimeError(f'readout: value has {n} memory elements, affinity {affinity.shape[0]}')
    hw = affinity.shape[1]
    out = torch.empty((rows, hw), dtype=torch.float32, device=value.device)
    if rows >= 64 and hw <= 256 and n >= 256 and v_ld % 4 == 0:
        # consolidation-sized products go through the tcgen05 split-K GEMM (csrc/keyproj.cu, tc_readout_dense)
        ws = torch.empty(int(N.lib.vosmem_readout_dense_tc_workspace_bytes(rows, n, hw)), dtype=torch.uint8, device=value.device)
        check(N.lib.vosmem_readout_dense_tc(value.data_ptr(), v_ld, affinity.data_ptr(), a_ld, rows, n, hw, out.data_ptr(), hw,
                                            ws.data_ptr(), ws.numel(), _stream()), 'vosmem_readout_dense_tc')
        return out
    check(N.lib.vosmem_readout_dense(value.data_ptr(), v_ld, affinity.data_ptr(), a_ld, rows, n, hw, out.data_ptr(), hw,
                                     _stream()), 'vosmem_readout_dense')
    return out
