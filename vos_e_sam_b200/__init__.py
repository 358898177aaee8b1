"""vos_e_sam_b200 -- B200-native (sm_100a) space-time memory readout for the VOS-E-SAM / XMem tracker.

Drop-in replacements for the reference's hot path, backed by hand-written CUDA kernels behind a C ABI
(include/vosmem.h -> vos_e_sam_b200/lib/libvosmem.so):

    vos_e_sam_b200.memory_util       <-> tracker/model/memory_util.py
    vos_e_sam_b200.kv_memory_store   <-> tracker/inference/kv_memory_store.py
    vos_e_sam_b200.memory_manager    <-> tracker/inference/memory_manager.py
    vos_e_sam_b200.key_projection    <-> tracker/model/modules.py KeyProjection (the producer of the query operands)
    vos_e_sam_b200.dropin.install()  makes the reference's InferenceCore pick them up unchanged
    vos_e_sam_b200.sharded           N-sharded long-term memory / data-parallel sequences over NCCL

Importing the package loads the shared library and fails loudly when it has not been built
(python -m vos_e_sam_b200.build).  There is no CPU or PyTorch fallback.
"""
import importlib

_LAZY = {
    'MemoryManager': 'memory_manager', 'match_memory_batch': 'memory_manager', 'KeyValueMemoryStore': 'kv_memory_store',
    'KeyProjection': 'key_projection',
    'get_similarity': 'memory_util', 'do_softmax': 'memory_util', 'get_affinity': 'memory_util', 'readout': 'memory_util',
}
__all__ = sorted(_LAZY)


def __getattr__(name):
    # Resolved on first use so that `vos_e_sam_b200.build` can (re)build the library before anything loads it;
    # every one of these modules imports `_native`, which raises if libvosmem.so is missing or stale.
    if name in _LAZY:
        return getattr(importlib.import_module(f'{__name__}.{_LAZY[name]}'), name)
    if name in ('ops', '_native', 'memory_manager', 'kv_memory_store', 'memory_util', 'dropin', 'sharded', 'build', 'key_projection'):
        return importlib.import_module(f'{__name__}.{name}')
    raise AttributeError(f'module {__name__!r} has no attribute {name!r}')
