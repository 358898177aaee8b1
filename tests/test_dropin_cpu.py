"""CPU: dropin.install() re-routes the reference's import points to this package (no compute; needs the reference
tree, which only exists in the authoring container -- skipped elsewhere)."""
import os
import sys

import pytest

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'tracker')), reason='reference tree not mounted')


@pytest.fixture
def ref_paths():
    added = [REF, os.path.join(REF, 'tracker')]
    sys.path[:0] = added
    before = set(sys.modules)
    yield
    for p in added:
        sys.path.remove(p)
    for name in set(sys.modules) - before:
        if name.split('.')[0] in ('tracker', 'model', 'inference', 'util'):
            sys.modules.pop(name, None)


def test_install_routes_inference_core_to_the_b200_manager(ref_paths):
    from vos_e_sam_b200 import build
    build.build()
    import vos_e_sam_b200 as vos
    from vos_e_sam_b200 import dropin
    import tracker.inference.inference_core as ic            # imported BEFORE install: names get re-bound
    ref_manager = ic.MemoryManager
    assert ref_manager.__module__ == 'tracker.inference.memory_manager'
    dropin.install()
    try:
        assert ic.MemoryManager is vos.MemoryManager
        import tracker.inference.memory_manager as mm        # module alias
        assert mm.MemoryManager is vos.MemoryManager
        import model.memory_util as mu
        assert mu.get_similarity is vos.get_similarity and mu.do_softmax is vos.do_softmax

        class StubNetwork:                                   # InferenceCore only stores it at construction
            pass
        cfg = dict(mem_every=5, deep_update_every=-1, enable_long_term=True, enable_long_term_count_usage=True,
                   hidden_dim=64, top_k=30, max_mid_term_frames=10, min_mid_term_frames=5, num_prototypes=128,
                   max_long_term_elements=10000)
        core = ic.InferenceCore(StubNetwork(), cfg)
        assert isinstance(core.memory, vos.MemoryManager)
        core.clear_memory()                                  # tools/runner.py:61 path: a fresh manager per video
        assert isinstance(core.memory, vos.MemoryManager) and core.memory.work_mem.size == 0
        # same attribute surface InferenceCore / callers touch
        for attr in ('match_memory', 'add_memory', 'create_hidden_state', 'set_hidden', 'get_hidden', 'update_config',
                     'work_mem', 'long_mem', 'hidden', 'CK', 'CV', 'H', 'W'):
            assert hasattr(core.memory, attr)
    finally:
        dropin.uninstall()
    assert ic.MemoryManager is ref_manager
    assert 'tracker.inference.memory_manager' not in sys.modules or \
        sys.modules['tracker.inference.memory_manager'].MemoryManager is ref_manager


def test_signatures_match_the_reference(ref_paths):
    """Same parameter names, order and defaults as the reference's public functions / methods."""
    import inspect
    import vos_e_sam_b200 as vos
    from vos_e_sam_b200 import memory_util as ours_util
    import model.memory_util as ref_util
    from tracker.inference.memory_manager import MemoryManager as RefManager
    from tracker.inference.kv_memory_store import KeyValueMemoryStore as RefStore
    for name in ('get_similarity', 'do_softmax', 'get_affinity', 'readout'):
        ours, ref = inspect.signature(getattr(ours_util, name)), inspect.signature(getattr(ref_util, name))
        assert list(ours.parameters) == list(ref.parameters), name
        assert [p.default for p in ours.parameters.values()] == [p.default for p in ref.parameters.values()], name
    for name in ('match_memory', 'add_memory', 'create_hidden_state', 'set_hidden', 'get_hidden', 'update_config',
                 'compress_features', 'consolidation', '_readout'):
        ours = list(inspect.signature(getattr(vos.MemoryManager, name)).parameters)
        ref = list(inspect.signature(getattr(RefManager, name)).parameters)
        assert ours == ref, name
    for name in ('add', 'update_usage', 'sieve_by_range', 'remove_obsolete_features', 'get_usage', 'get_all_sliced',
                 'get_v_size', 'engaged'):
        ours = list(inspect.signature(getattr(vos.KeyValueMemoryStore, name)).parameters)
        ref = list(inspect.signature(getattr(RefStore, name)).parameters)
        assert ours == ref, name
    for prop in ('size', 'num_groups', 'key', 'value', 'shrinkage', 'selection'):
        assert isinstance(getattr(vos.KeyValueMemoryStore, prop), property) and isinstance(getattr(RefStore, prop), property)


def test_key_projection_is_interchangeable_with_the_reference_module(ref_paths):
    """vos_e_sam_b200.KeyProjection vs tracker/model/modules.py:194-211: same constructor and forward parameters, the
    same parameter names and shapes -- a state_dict of one loads into the other (how an XMem checkpoint reaches it)."""
    import inspect
    import torch
    import vos_e_sam_b200 as vos
    from model.modules import KeyProjection as RefKeyProjection
    for fn in ('__init__', 'forward'):
        ours = list(inspect.signature(getattr(vos.KeyProjection, fn)).parameters)
        ref = list(inspect.signature(getattr(RefKeyProjection, fn)).parameters)
        assert ours == ref, fn
    ref, ours = RefKeyProjection(64, 64), vos.KeyProjection(64, 64)
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    ours.load_state_dict(ref.state_dict())
    ref.load_state_dict(ours.state_dict())
    assert torch.equal(ours.key_proj.weight, ref.key_proj.weight)
    with pytest.raises(RuntimeError, match='CUDA'):         # no CPU path
        ours(torch.zeros(1, 64, 4, 4), True, True)
