"""CPU: the C-ABI library builds, loads and exports every symbol include/vosmem.h declares.

No compute call is made here (no GPU in the authoring container); host-only helpers are exercised.
"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'vosmem.h')


@pytest.fixture(scope='module')
def native():
    from vos_e_sam_b200 import build
    build.build()
    from vos_e_sam_b200 import _native
    return _native


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(vosmem_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for must in ('vosmem_match', 'vosmem_select_topk', 'vosmem_merge_topk', 'vosmem_softmax_readout', 'vosmem_pack_keys',
                 'vosmem_pack_values', 'vosmem_similarity_dense', 'vosmem_softmax_dense', 'vosmem_readout_dense'):
        assert must in names


def test_library_exports_every_declared_symbol(native):
    lib = ctypes.CDLL(native.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f'{name} declared in include/vosmem.h but not exported by libvosmem.so'


def test_python_binding_covers_the_header(native):
    assert sorted(native.SIGNATURES) == declared_functions()


def test_struct_layouts_match_the_header(native):
    # sizes follow from the C declaration order (int / pointer / int64 with natural alignment)
    assert ctypes.sizeof(native.Segment) == 48
    assert ctypes.sizeof(native.SelectDesc) == 16 + 16 + 8 + 2 * 48 + 8 + 8 + 8 + 8
    assert ctypes.sizeof(native.ValueSegment) == 48
    assert ctypes.sizeof(native.AppendGroup) == 64 and ctypes.sizeof(native.AppendDesc) == 128 + 8 * 64
    assert ctypes.sizeof(native.ReadoutDesc) == 24 + 2 * 48 + 24 + 16


def test_host_side_layout_helpers(native):
    lib = native.lib
    assert lib.vosmem_abi_version() == 3
    # 64 keys per tile; 34 sixteen-byte chunks per key (16 hi + 16 lo + 2 tail)
    assert lib.vosmem_key_image_bytes(64, 64) == 64 * 34 * 16
    assert lib.vosmem_key_image_bytes(64, 65) == 2 * 64 * 34 * 16
    assert lib.vosmem_key_image_bytes(64, 0) == 0
    assert lib.vosmem_key_image_bytes(32, 64) == 0            # no image format for other CK
    small, big = lib.vosmem_workspace_bytes(64, 1620, 16200), lib.vosmem_workspace_bytes(64, 8160, 100000)
    assert 0 < small < big < 256 * 2 ** 20
    assert lib.vosmem_status_string(0) == b'ok'
    assert lib.vosmem_status_string(-22) == b'invalid argument'


def test_argument_errors_do_not_need_a_gpu(native):
    # validation happens before any CUDA call, so these run on the CPU box
    lib = native.lib
    d = native.SelectDesc()
    assert lib.vosmem_select_topk(ctypes.byref(d), None, None, None) == -22
    assert b'null output' in lib.vosmem_last_error()
    r = native.ReadoutDesc()
    assert lib.vosmem_match(ctypes.byref(d), ctypes.byref(r), None, None, None) == -22
    assert b'CK' in lib.vosmem_last_error()
    assert lib.vosmem_pack_keys(None, 0, None, 32, 0, 0, None, 0, None) == -22
    assert lib.vosmem_merge_topk(None, None, 1, 1, 30, None, None, None) == -22
    with pytest.raises(native.VosmemError):
        native.check(-22, 'demo')
