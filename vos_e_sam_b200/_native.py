"""ctypes binding of libvosmem.so (the C ABI declared in include/vosmem.h).

There is no CPU or PyTorch fallback behind this module: if the shared library is missing the import
fails loudly and tells the caller how to build it.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'lib', 'libvosmem.so')

OK = 0
ABI_VERSION = 3
F32, BF16 = 0, 1
PATH_AUTO, PATH_SIMT, PATH_TCGEN05 = 0, 1, 2
MAX_TOPK = 32
MAX_BATCH = 12
KEY_TILE = 64
QUERY_TILE = 128

vp = C.c_void_p
i64 = C.c_int64


class Segment(C.Structure):
    _fields_ = [('key', vp), ('key_ld', i64), ('shrinkage', vp), ('key_image', vp), ('begin', i64), ('end', i64)]


class SelectDesc(C.Structure):
    _fields_ = [('ck', C.c_int), ('hw', C.c_int), ('top_k', C.c_int), ('query_key', vp), ('query_selection', vp),
                ('n_segments', C.c_int), ('seg', Segment * 2), ('index_base', i64), ('path', C.c_int),
                ('workspace', vp), ('workspace_bytes', i64)]


class ValueSegment(C.Structure):
    _fields_ = [('shadow', vp), ('shadow_ld', i64), ('first', i64), ('count', i64), ('use_count', vp), ('life_count', vp)]


class ReadoutDesc(C.Structure):
    _fields_ = [('hw', C.c_int), ('top_k', C.c_int), ('rows', C.c_int), ('value_dtype', C.c_int),
                ('n_segments', C.c_int), ('seg', ValueSegment * 2), ('out', vp), ('out_ld', i64), ('out_weight', vp),
                ('out_group_rows', C.c_int), ('out_group_stride', i64)]


MAX_RANKS = 16
EXCH_K = 32


class PushDesc(C.Structure):
    _fields_ = [('world', C.c_int), ('rank', C.c_int), ('per', C.c_int), ('index_base', i64), ('dst', vp * MAX_RANKS),
                ('flag', vp * MAX_RANKS), ('seq', C.c_uint32), ('ticket', vp),
                ('rank_pub', vp * MAX_RANKS), ('seg0_len', i64), ('index_base1', i64)]


MAX_GROUPS = 8


class AppendGroup(C.Structure):
    _fields_ = [('value', vp), ('value_ld', i64), ('rows', C.c_int), ('ref', vp), ('ref_ld', i64), ('shadow', vp),
                ('shadow_ld', i64), ('n', i64)]


class AppendDesc(C.Structure):
    _fields_ = [('ck', C.c_int), ('m', C.c_int), ('n', i64), ('key', vp), ('key_ld', i64), ('shrinkage', vp),
                ('selection', vp), ('selection_ld', i64), ('bank_key', vp), ('bank_ld', i64), ('bank_shrinkage', vp),
                ('bank_selection', vp), ('bank_use', vp), ('bank_life', vp), ('key_image', vp), ('capacity', i64),
                ('value_dtype', C.c_int), ('n_groups', C.c_int), ('group', AppendGroup * MAX_GROUPS)]


class ExchangeDesc(C.Structure):
    _fields_ = [('lists', vp), ('n_lists', C.c_int), ('list_stride', i64), ('first_entry', i64), ('flags', vp),
                ('seq', C.c_uint32), ('status', vp)]


# name -> (restype, argtypes); mirrors include/vosmem.h one to one (tests check the two agree)
SIGNATURES = {
    'vosmem_abi_version': (C.c_int, []),
    'vosmem_last_error': (C.c_char_p, []),
    'vosmem_status_string': (C.c_char_p, [C.c_int]),
    'vosmem_key_image_bytes': (i64, [C.c_int, i64]),
    'vosmem_workspace_bytes': (i64, [C.c_int, C.c_int, i64]),
    'vosmem_query_image_bytes': (i64, [C.c_int, C.c_int]),
    'vosmem_softmax_dense_scratch_bytes': (i64, [i64, C.c_int]),
    'vosmem_softmax_dense_ws': (C.c_int, [vp, i64, i64, C.c_int, C.c_int, vp, i64, vp, vp, i64, vp]),
    'vosmem_readout_dense_tc_workspace_bytes': (i64, [C.c_int, i64, C.c_int]),
    'vosmem_readout_dense_tc': (C.c_int, [vp, i64, vp, i64, C.c_int, i64, C.c_int, vp, i64, vp, i64, vp]),
    'vosmem_keyproj_weight_bytes': (i64, [C.c_int, C.c_int]),
    'vosmem_keyproj_pack_weights': (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, vp]),
    'vosmem_keyproj_workspace_bytes': (i64, [C.c_int, C.c_int, C.c_int, C.c_int]),
    'vosmem_keyproj_forward': (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]),
    'vosmem_workspace_init': (C.c_int, [vp, i64, vp]),
    'vosmem_workspace_status': (C.c_int, [vp, vp]),
    'vosmem_store_append': (C.c_int, [vp, vp]),
    'vosmem_pack_keys': (C.c_int, [vp, i64, vp, C.c_int, i64, i64, vp, i64, vp]),
    'vosmem_pack_values': (C.c_int, [vp, i64, C.c_int, i64, i64, vp, i64, i64, C.c_int, vp]),
    'vosmem_select_topk': (C.c_int, [C.POINTER(SelectDesc), vp, vp, vp]),
    'vosmem_merge_topk': (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    'vosmem_merge_topk_ptrs': (C.c_int, [C.POINTER(vp), C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    'vosmem_softmax_readout': (C.c_int, [C.POINTER(ReadoutDesc), vp, vp, vp]),
    'vosmem_match': (C.c_int, [C.POINTER(SelectDesc), C.POINTER(ReadoutDesc), vp, vp, vp]),
    'vosmem_match_batch': (C.c_int, [C.POINTER(SelectDesc), C.POINTER(ReadoutDesc), C.c_int, vp]),
    'vosmem_age': (C.c_int, [vp, i64, vp]),
    'vosmem_select_push': (C.c_int, [C.POINTER(SelectDesc), C.POINTER(PushDesc), vp]),
    'vosmem_exchange_readout': (C.c_int, [C.POINTER(ReadoutDesc), C.POINTER(ExchangeDesc), vp]),
    'vosmem_push_slice': (C.c_int, [vp, i64, C.c_int, C.c_int, C.POINTER(vp), i64, C.POINTER(vp), C.c_int, C.c_uint32, vp, vp]),
    'vosmem_wait_flags': (C.c_int, [vp, C.c_uint32, C.c_uint32, vp, vp]),
    'vosmem_similarity_dense': (C.c_int, [vp, i64, vp, vp, vp, C.c_int, i64, C.c_int, vp, vp]),
    'vosmem_softmax_dense': (C.c_int, [vp, i64, i64, C.c_int, C.c_int, vp, i64, vp, vp]),
    'vosmem_readout_dense': (C.c_int, [vp, i64, vp, i64, C.c_int, i64, C.c_int, vp, i64, vp]),
    'vosmem_debug_umma_tile': (C.c_int, [vp, vp, vp, vp]),
    'vosmem_debug_pack_query': (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp]),
    'vosmem_debug_set_timing_buffer': (C.c_int, [vp]),
    'vosmem_debug_set_stage_events': (C.c_int, [vp, vp, vp, vp]),
}


class NativeLibraryMissing(ImportError):
    pass


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f'{LIB_PATH} not found: build the CUDA extension first (python -m vos_e_sam_b200.build, or '
            f'__graft_entry__.build()).  vos_e_sam_b200 has no CPU / PyTorch fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    if lib.vosmem_abi_version() != ABI_VERSION:
        raise NativeLibraryMissing(f'{LIB_PATH}: ABI version {lib.vosmem_abi_version()} != {ABI_VERSION}, rebuild')
    return lib


lib = _load()


class VosmemError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        msg = lib.vosmem_last_error().decode(errors='replace')
        kind = lib.vosmem_status_string(status).decode()
        super().__init__(f'{where}: {kind} ({status}): {msg}')


def check(status: int, where: str) -> None:
    if status != OK:
        raise VosmemError(status, where)
