"""Generate golden vectors by running the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the authoring container (where the reference is
mounted read-only at /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

The GPU box has no /root/reference, so the vectors are committed.  Three files:

  util_cases.npz       memory_util.get_similarity / do_softmax / readout on small seeded inputs
  match_cases.npz      MemoryManager.match_memory on hand-built stores: working-only, multi-group,
                       long-term engaged, long-term with a group that has no prototypes, usage off
  lifecycle_*.npz      every MemoryManager call made by the reference InferenceCore.step driven by a
                       stub network over a short video (add_memory, compression into prototypes,
                       least-used eviction, a second object group appearing mid-video)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF = '/root/reference'
sys.path[:0] = [REF, os.path.join(REF, 'tracker')]

from tracker.inference.memory_manager import MemoryManager      # noqa: E402
from tracker.inference.inference_core import InferenceCore      # noqa: E402
import model.memory_util as ref_util                            # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
CK = 64


def rnd(gen, *shape):
    return torch.randn(*shape, generator=gen)


def synth_keys(gen, n):
    """keys ~ N(0,1); shrinkage = 1 + N(0,1)^2 (modules.py:208); selection = sigmoid(N(0,1)) (modules.py:209)."""
    return rnd(gen, 1, CK, n), 1 + rnd(gen, 1, 1, n) ** 2, torch.sigmoid(rnd(gen, 1, CK, n))


# ---------------------------------------------------------------------------
def util_cases():
    g = torch.Generator().manual_seed(101)
    out = {}
    n, hw, cv, k = 200, 48, 16, 30
    mk, ms, _ = synth_keys(g, n)
    qk, _, qe = synth_keys(g, hw)
    out['mk'], out['ms'], out['qk'], out['qe'] = mk, ms, qk, qe
    out['sim_aniso'] = ref_util.get_similarity(mk, ms, qk, qe)
    out['sim_iso'] = ref_util.get_similarity(mk, ms, qk, None)
    out['sim_noshrink'] = ref_util.get_similarity(mk, None, qk, qe)
    out['sim_plain'] = ref_util.get_similarity(mk, None, qk, None)
    sim = out['sim_aniso']
    aff, usage = ref_util.do_softmax(sim.clone(), top_k=k, inplace=False, return_usage=True)
    out['aff_topk'], out['usage_topk'] = aff, usage
    out['aff_topk_inplace'] = ref_util.do_softmax(sim.clone(), top_k=k, inplace=True)
    daff, dusage = ref_util.do_softmax(sim.clone(), top_k=None, return_usage=True)
    out['aff_dense'], out['usage_dense'] = daff, dusage
    out['aff_get_affinity'] = ref_util.get_affinity(mk, ms, qk, qe)
    # readout: mv B x CV x T x H x W with T*H*W == n  (memory_util.py:73-80)
    t, h, w = 5, 5, 8
    mv = rnd(g, 1, cv, t, h, w)
    aff_r = ref_util.do_softmax(ref_util.get_similarity(mk, ms, rnd(g, 1, CK, h, w), torch.sigmoid(rnd(g, 1, CK, h, w))),
                                top_k=k)
    out['mv'], out['aff_for_readout'] = mv, aff_r
    out['readout'] = ref_util.readout(aff_r, mv)
    out['top_k'] = torch.tensor(k)
    np.savez_compressed(os.path.join(OUT, 'util_cases.npz'), **{a: b.numpy() for a, b in out.items()})


# ---------------------------------------------------------------------------
def base_config(**over):
    cfg = dict(hidden_dim=8, top_k=30, enable_long_term=True, enable_long_term_count_usage=True,
               max_mid_term_frames=10, min_mid_term_frames=5, num_prototypes=128,
               max_long_term_elements=10000, mem_every=5, deep_update_every=-1)
    cfg.update(over)
    return cfg


def build_manager(gen, cfg, hw, cv, work_frames, groups, long_n=0, long_groups=0):
    """Hand-build a reference MemoryManager: ``groups`` = list of (object ids, first frame index)."""
    mm = MemoryManager(cfg)
    h, w = hw
    first_seen = {}
    for objs, t0 in groups:
        for o in objs:
            first_seen[o] = t0
    for t in range(work_frames):
        present = sorted(o for o, t0 in first_seen.items() if t0 <= t)
        n_all = max(present)
        k, s, e = rnd(gen, 1, CK, h, w), 1 + rnd(gen, 1, 1, h, w) ** 2, torch.sigmoid(rnd(gen, 1, CK, h, w))
        v = rnd(gen, 1, n_all, cv, h, w)
        mm.add_memory(k, s, v, present, selection=e if cfg['enable_long_term'] else None)
    if long_n:
        pk, ps, _ = synth_keys(gen, long_n)
        pv = []
        for gi in range(long_groups):
            n_obj = mm.work_mem.value[gi].shape[0]
            # later groups own only a suffix of the long-term keys
            n_g = long_n if gi == 0 else long_n // (gi + 1)
            pv.append(rnd(gen, n_obj, cv, n_g))
        mm.long_mem.add(pk, pv, ps, selection=None, objects=None)
        if cfg['enable_long_term_count_usage']:
            mm.long_mem.use_count += torch.rand(1, 1, long_n, generator=gen)
            mm.long_mem.life_count += 3
    if cfg['enable_long_term']:
        mm.work_mem.use_count += torch.rand(mm.work_mem.use_count.shape, generator=gen)
        mm.work_mem.life_count += 2
    return mm


def dump_store(prefix, store, out):
    out[prefix + 'key'] = store.key
    out[prefix + 'shrinkage'] = store.shrinkage
    if store.selection is not None:
        out[prefix + 'selection'] = store.selection
    for gi, gv in enumerate(store.value):
        out[f'{prefix}value{gi}'] = gv
    out[prefix + 'num_groups'] = torch.tensor(store.num_groups)
    if store.count_usage:
        out[prefix + 'use_count'] = store.use_count.clone()
        out[prefix + 'life_count'] = store.life_count.clone()


def match_cases():
    specs = {
        # name: (cfg overrides, hw, cv, work_frames, groups, long_n, long_groups)
        'w1': ({}, (6, 8), 32, 3, [([1, 2], 0)], 0, 0),
        'w2': ({}, (6, 8), 32, 4, [([1], 0), ([2, 3], 2)], 0, 0),
        'l1': ({}, (6, 8), 32, 3, [([1, 2], 0)], 160, 1),
        'l3': ({}, (5, 7), 16, 5, [([1], 0), ([2], 1), ([3, 4], 3)], 150, 2),
        'nousage': (dict(enable_long_term=False, enable_long_term_count_usage=False), (6, 8), 32, 3, [([1, 2], 0)], 0, 0),
        'nolongusage': (dict(enable_long_term_count_usage=False), (6, 8), 32, 3, [([1], 0)], 100, 1),
        'k5': (dict(top_k=5), (4, 4), 8, 2, [([1], 0)], 0, 0),
    }
    out = {}
    names = []
    for i, (name, (over, hw, cv, wf, groups, long_n, long_groups)) in enumerate(specs.items()):
        g = torch.Generator().manual_seed(500 + i)
        cfg = base_config(**over)
        mm = build_manager(g, cfg, hw, cv, wf, groups, long_n, long_groups)
        p = name + '/'
        dump_store(p + 'work_', mm.work_mem, out)
        if cfg['enable_long_term'] and mm.long_mem.engaged():
            dump_store(p + 'long_', mm.long_mem, out)
        qk, qe = rnd(g, 1, CK, *hw), torch.sigmoid(rnd(g, 1, CK, *hw))
        out[p + 'qk'], out[p + 'qe'] = qk, qe
        out[p + 'readout'] = mm.match_memory(qk, qe)
        if mm.work_mem.count_usage:
            out[p + 'work_use_after'] = mm.work_mem.use_count.clone()
            out[p + 'work_life_after'] = mm.work_mem.life_count.clone()
        if cfg['enable_long_term'] and mm.long_mem.engaged() and mm.long_mem.count_usage:
            out[p + 'long_use_after'] = mm.long_mem.use_count.clone()
            out[p + 'long_life_after'] = mm.long_mem.life_count.clone()
        out[p + 'cfg'] = torch.tensor([cfg['top_k'], int(cfg['enable_long_term']),
                                       int(cfg['enable_long_term_count_usage']), cv])
        names.append(name)
    arrays = {a: b.numpy() for a, b in out.items()}
    arrays['names'] = np.array(names)
    np.savez_compressed(os.path.join(OUT, 'match_cases.npz'), **arrays)


# ---------------------------------------------------------------------------
class StubNetwork:
    """Deterministic stand-in for XMem (network.py:40,72,107): random features, no conv nets."""

    def __init__(self, seed, cv):
        self.gen = torch.Generator().manual_seed(seed)
        self.cv = cv

    def encode_key(self, image, need_ek=True, need_sk=True):
        h, w = image.shape[-2] // 16, image.shape[-1] // 16
        key = rnd(self.gen, 1, CK, h, w)
        shrinkage = 1 + rnd(self.gen, 1, 1, h, w) ** 2 if need_sk else None
        selection = torch.sigmoid(rnd(self.gen, 1, CK, h, w)) if need_ek else None
        return key, shrinkage, selection, None, None, None

    def segment(self, feats, memory_readout, hidden, h_out=True, strip_bg=False):
        n = memory_readout.shape[1]
        H, W = memory_readout.shape[-2] * 16, memory_readout.shape[-1] * 16
        logits = rnd(self.gen, 1, n + 1, H, W) + memory_readout.mean()
        prob = torch.softmax(logits, dim=1)
        return hidden, logits, prob

    def encode_value(self, image, f16, hidden, masks, is_deep_update=True):
        n = masks.shape[1]
        h, w = image.shape[-2] // 16, image.shape[-1] // 16
        return rnd(self.gen, 1, n, self.cv, h, w), hidden


def lifecycle(name, cfg, frames, hw, cv, new_object_at=None):
    """Drive the reference InferenceCore and log every MemoryManager.match_memory / add_memory call."""
    h, w = hw
    H, W = h * 16, w * 16
    net = StubNetwork(seed=900 + len(name), cv=cv)
    core = InferenceCore(net, cfg)
    g = torch.Generator().manual_seed(77)
    log = {}
    events = []

    def arm(mm):
        ref_match, ref_add = mm.match_memory, mm.add_memory

        def match(query_key, selection):
            res = ref_match(query_key, selection)
            i = len(events)
            log[f'{i}/qk'], log[f'{i}/qe'], log[f'{i}/readout'] = query_key.clone(), selection.clone(), res.clone()
            log[f'{i}/sizes'] = torch.tensor([mm.work_mem.size, mm.long_mem.size])
            events.append('match')
            return res

        def add(key, shrinkage, value, objects, selection=None):
            i = len(events)
            log[f'{i}/key'], log[f'{i}/shrinkage'], log[f'{i}/value'] = key.clone(), shrinkage.clone(), value.clone()
            log[f'{i}/selection'] = selection.clone()
            log[f'{i}/objects'] = torch.tensor(list(objects))
            ref_add(key, shrinkage, value, objects, selection=selection)
            log[f'{i}/sizes'] = torch.tensor([mm.work_mem.size, mm.long_mem.size])
            events.append('add')

        mm.match_memory, mm.add_memory = match, add

    arm(core.memory)
    labels = [1, 2]
    core.set_all_labels(labels)
    for t in range(frames):
        image = rnd(g, 3, H, W)
        if t == 0:
            mask = (torch.rand(len(labels), H, W, generator=g) > 0.7).float()
            core.step(image, mask, labels)
        elif new_object_at is not None and t == new_object_at:
            labels = labels + [3]
            core.set_all_labels(labels)
            mask = (torch.rand(len(labels), H, W, generator=g) > 0.8).float()
            core.step(image, mask, [3])
        else:
            core.step(image)
    mm = core.memory
    dump_store('final/work_', mm.work_mem, log)
    if mm.long_mem.engaged():
        dump_store('final/long_', mm.long_mem, log)
    arrays = {a: b.numpy() for a, b in log.items()}
    arrays['events'] = np.array(events)
    arrays['cfg_top_k'] = np.array(cfg['top_k'])
    arrays['cfg_max_long'] = np.array(cfg['max_long_term_elements'])
    arrays['cfg_num_prototypes'] = np.array(cfg['num_prototypes'])
    arrays['cfg_min_mid'] = np.array(cfg['min_mid_term_frames'])
    arrays['cfg_max_mid'] = np.array(cfg['max_mid_term_frames'])
    np.savez_compressed(os.path.join(OUT, f'lifecycle_{name}.npz'), **arrays)
    print(name, 'events', len(events), 'final sizes', mm.work_mem.size, mm.long_mem.size,
          'groups', mm.work_mem.num_groups)


def keyproj_cases():
    """The reference's KeyProjection module (tracker/model/modules.py:194-211) on seeded parameters and inputs.  Only the
    outputs are stored: tests/synth.keyproj_params and the input are regenerated from the same seeds by the tests."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tests import synth
    from model.modules import KeyProjection
    out = {}
    for name, in_dim, h, w in (('small', 64, 5, 7), ('xmem', 1024, 12, 20)):
        g = torch.Generator().manual_seed(500 + in_dim)
        prm = synth.keyproj_params(g, in_dim)
        x = rnd(g, 1, in_dim, h, w)
        mod = KeyProjection(in_dim, CK)
        with torch.no_grad():
            mod.key_proj.weight.copy_(prm['key_w']); mod.key_proj.bias.copy_(prm['key_b'])
            mod.d_proj.weight.copy_(prm['d_w']); mod.d_proj.bias.copy_(prm['d_b'])
            mod.e_proj.weight.copy_(prm['e_w']); mod.e_proj.bias.copy_(prm['e_b'])
            key, shrinkage, selection = mod(x, True, True)
            key_only = mod(x, False, False)
        assert key_only[1] is None and key_only[2] is None and torch.equal(key_only[0], key)
        out[f'{name}/shape'] = torch.tensor([in_dim, h, w])
        out[f'{name}/key'], out[f'{name}/shrinkage'], out[f'{name}/selection'] = key, shrinkage, selection
    np.savez_compressed(os.path.join(OUT, 'keyproj_cases.npz'), **{a: b.numpy() for a, b in out.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)          # bit-stable reductions for the committed vectors
    if '--keyproj-only' in sys.argv:
        keyproj_cases()
        return
    util_cases()
    keyproj_cases()
    match_cases()
    # single group, small long-term budget -> compression + least-used eviction
    lifecycle('evict', base_config(top_k=8, mem_every=2, num_prototypes=32, max_long_term_elements=80),
              frames=56, hw=(4, 6), cv=8)
    # a third object appears at frame 9 -> second object group; no eviction (unsupported with >1 group)
    lifecycle('groups', base_config(top_k=8, mem_every=2, num_prototypes=32, max_long_term_elements=4000),
              frames=40, hw=(4, 6), cv=8, new_object_at=9)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, 'KiB')


if __name__ == '__main__':
    main()
