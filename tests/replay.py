"""Helpers shared by the CPU (oracle) and GPU (product) golden tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def t(a, device='cpu'):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def lifecycle_config(z):
    return dict(hidden_dim=8, top_k=int(z['cfg_top_k']), enable_long_term=True,
                enable_long_term_count_usage=True, max_mid_term_frames=int(z['cfg_max_mid']),
                min_mid_term_frames=int(z['cfg_min_mid']), num_prototypes=int(z['cfg_num_prototypes']),
                max_long_term_elements=int(z['cfg_max_long']), mem_every=2, deep_update_every=-1)


def replay_lifecycle(z, manager, device='cpu', on_match=None):
    """Replay every recorded MemoryManager call; returns list of (event index, readout, expected)."""
    results = []
    for i, ev in enumerate(z['events']):
        if ev == 'match':
            got = manager.match_memory(t(z[f'{i}/qk'], device), t(z[f'{i}/qe'], device))
            results.append((i, got, t(z[f'{i}/readout'])))
            if on_match is not None:
                on_match(i, got)
        else:
            manager.add_memory(t(z[f'{i}/key'], device), t(z[f'{i}/shrinkage'], device),
                               t(z[f'{i}/value'], device), [int(o) for o in z[f'{i}/objects']],
                               selection=t(z[f'{i}/selection'], device))
        sizes = z[f'{i}/sizes']
        long_size = manager.long_mem.size if manager.long_mem is not None else 0
        assert (manager.work_mem.size, long_size) == (int(sizes[0]), int(sizes[1])), (i, ev)
    return results
