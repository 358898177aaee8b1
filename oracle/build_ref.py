"""Stage the UNMODIFIED reference modules of the hot path under oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

    python oracle/build_ref.py        # in the authoring container, where /root/reference is mounted

The reference is pure Python (no setup.py, nothing to compile), so its "build" is a byte-for-byte copy of the import
closure of ``tracker/inference/inference_core.py`` -- the memory readout (memory_util / memory_manager /
kv_memory_store), InferenceCore, and the model files those import -- with a sha256 manifest.  Nothing is edited.
``oracle/_ref/`` is git-ignored (reference sources never enter the history) but travels to the GPU box with the
snapshot, like the built .so, where it serves as

  * the CPU baseline / ``bench.py --impl reference`` arm (``cpu_baseline.kind == "reference"``),
  * the secondary baseline: the reference's own torch op sequence on CUDA tensors (``gpu_eager_baseline``),
  * the reference side of the GPU drop-in test (reference InferenceCore.step + dropin.install()).

``load()`` imports the staged copy under its original module names (``tracker.inference.memory_manager``,
``model.memory_util`` ...).  Only tests/, smoke() and bench.py's baseline legs may call it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref')
SRC = '/root/reference'
FILES = [
    'tracker/inference/__init__.py', 'tracker/inference/inference_core.py', 'tracker/inference/memory_manager.py',
    'tracker/inference/kv_memory_store.py',
    'tracker/model/__init__.py', 'tracker/model/memory_util.py', 'tracker/model/network.py', 'tracker/model/aggregate.py',
    'tracker/model/modules.py', 'tracker/model/group_modules.py', 'tracker/model/resnet.py', 'tracker/model/cbam.py',
    'tracker/util/__init__.py', 'tracker/util/tensor_util.py',
]


def _sha(path: str) -> str:
    with open(path, 'rb') as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(src: str = SRC) -> bool:
    """Copy the file closure; returns False (and leaves an existing copy alone) when the reference is not mounted."""
    if not os.path.isdir(src):
        return False
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not os.path.exists(d) or _sha(d) != _sha(s):
            shutil.copyfile(s, d)
        manifest[rel] = _sha(d)
    with open(os.path.join(DEST, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': src, 'sha256': manifest}, f, indent=1)
    return True


def available() -> bool:
    return os.path.exists(os.path.join(DEST, 'MANIFEST.json'))


def load():
    """Import the staged reference; returns (MemoryManager class, memory_util module, InferenceCore class)."""
    if not available():
        raise ImportError('oracle/_ref is not staged: run python oracle/build_ref.py where /root/reference is mounted')
    for p in (os.path.join(DEST, 'tracker'), DEST):
        if p not in sys.path:
            sys.path.insert(0, p)
    from tracker.inference.memory_manager import MemoryManager
    from tracker.inference.inference_core import InferenceCore
    import model.memory_util as memory_util
    assert os.path.realpath(memory_util.__file__).startswith(os.path.realpath(DEST)), memory_util.__file__
    return MemoryManager, memory_util, InferenceCore


if __name__ == '__main__':
    ok = build()
    print('staged' if ok else 'reference not mounted', DEST)
