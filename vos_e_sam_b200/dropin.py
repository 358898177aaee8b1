"""Make the reference tracker use the B200 readout without editing it.

The reference has no plugin layer; the boundary is two import points:
``from tracker.inference.memory_manager import MemoryManager`` (tracker/inference/inference_core.py:1)
and ``from model.memory_util import *`` (tracker/inference/memory_manager.py:5, tracker/model/network.py:14).
``install()`` aliases those module names to this package and re-binds the names in consumers that were
already imported, so ``InferenceCore``, ``BaseTracker`` and the SAM refinement stage run unchanged.
"""
from __future__ import annotations

import sys

from . import kv_memory_store, memory_manager, memory_util

_MODULE_ALIASES = {
    'tracker.inference.memory_manager': memory_manager,
    'inference.memory_manager': memory_manager,
    'tracker.inference.kv_memory_store': kv_memory_store,
    'inference.kv_memory_store': kv_memory_store,
    'model.memory_util': memory_util,
    'tracker.model.memory_util': memory_util,
}
_CONSUMERS = {
    # module name -> names it imported from the hot-path modules
    'tracker.inference.inference_core': {'MemoryManager': memory_manager.MemoryManager},
    'inference.inference_core': {'MemoryManager': memory_manager.MemoryManager},
    'model.network': {n: getattr(memory_util, n) for n in memory_util.__all__},
    'tracker.model.network': {n: getattr(memory_util, n) for n in memory_util.__all__},
}
_saved = {}


def install() -> None:
    """Route the reference's imports of the memory readout to this package (idempotent)."""
    for name, mod in _MODULE_ALIASES.items():
        if name not in _saved:
            _saved[name] = sys.modules.get(name)
        sys.modules[name] = mod
        # `import a.b.c as x` binds getattr(a.b, 'c'), so an already-imported parent package is patched too
        parent, _, leaf = name.rpartition('.')
        pkg = sys.modules.get(parent)
        if pkg is not None and hasattr(pkg, leaf):
            _saved.setdefault((parent, leaf), getattr(pkg, leaf))
            setattr(pkg, leaf, mod)
    for name, bindings in _CONSUMERS.items():
        mod = sys.modules.get(name)
        if mod is None:
            continue
        for attr, obj in bindings.items():
            if hasattr(mod, attr):
                _saved.setdefault((name, attr), getattr(mod, attr))
                setattr(mod, attr, obj)


def uninstall() -> None:
    """Undo install(): restore the reference modules / names."""
    for key, old in list(_saved.items()):
        if isinstance(key, tuple):
            mod = sys.modules.get(key[0])
            if mod is not None:
                setattr(mod, key[1], old)
        elif old is None:
            sys.modules.pop(key, None)
        else:
            sys.modules[key] = old
    _saved.clear()
