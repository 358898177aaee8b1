"""vos_e_sam_b200 -- B200-native (sm_100a) space-time memory readout for the VOS-E-SAM / XMem tracker.

Drop-in replacements for the reference's hot path, backed by hand-written CUDA kernels behind a C ABI
(include/vosmem.h -> vos_e_sam_b200/lib/libvosmem.so):

    vos_e_sam_b200.memory_util       <-> tracker/model/memory_util.py
    vos_e_sam_b200.kv_memory_store   <-> tracker/inference/kv_memory_store.py
    vos_e_sam_b200.memory_manager    <-> tracker/inference/memory_manager.py
    vos_e_sam_b200.dropin.install()  makes the reference's InferenceCore pick them up unchanged
    vos_e_sam_b200.sharded           N-sharded long-term memory / data-parallel sequences over NCCL

Importing the package loads the shared library and fails loudly when it has not been built
(python -m vos_e_sam_b200.build).  There is no CPU or PyTorch fallback.
"""
from . import _native  # noqa: F401  (loads libvosmem.so or raises)
from .kv_memory_store import KeyValueMemoryStore
from .memory_manager import MemoryManager
from .memory_util import do_softmax, get_affinity, get_similarity, readout

__all__ = ['MemoryManager', 'KeyValueMemoryStore', 'get_similarity', 'do_softmax', 'get_affinity', 'readout']
