"""Where a working -> long-term consolidation (MemoryManager.compress_features, memory_manager.py:211-286) spends its
time on the GPU box: the steps called one by one, device-synchronised wall clock (us)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vos_e_sam_b200 as vos
from vos_e_sam_b200.memory_util import get_similarity, do_softmax
from tests import synth

dev = torch.device('cuda')
h, w, n_obj, cv = 30, 54, 5, 512
cfg = dict(hidden_dim=64, top_k=30, enable_long_term=True, enable_long_term_count_usage=True, max_mid_term_frames=10,
           min_mid_term_frames=5, num_prototypes=128, max_long_term_elements=10000)


def build():
    m = vos.MemoryManager(cfg)
    g = torch.Generator().manual_seed(3)
    for step in range(9):
        k, s, e = synth.keys(g, h * w)
        v = torch.randn(1, n_obj, cv, h, w, generator=g)
        m.add_memory(k.view(1, 64, h, w).to(dev), s.view(1, 1, h, w).to(dev), v.to(dev), list(range(1, n_obj + 1)),
                     selection=e.view(1, 64, h, w).to(dev))
        qk, qe = synth.query(g, h, w)
        m.match_memory(qk.to(dev), qe.to(dev))
    # the 10th frame, without the consolidation add_memory would trigger
    k, s, e = synth.keys(g, h * w)
    v = torch.randn(1, n_obj, cv, h, w, generator=g)
    m.work_mem.add(k.to(dev), v[0].flatten(2).to(dev), s.to(dev), e.to(dev), list(range(1, n_obj + 1)))
    return m


def timed(label, fn, acc):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    acc.setdefault(label, []).append((time.perf_counter() - t0) * 1e6)
    return out


acc = {}
for rep in range(4):
    m = build()
    hw = m.HW
    lo, hi = hw, -m.min_work_elements + hw
    work = m.work_mem
    cand_v = [gv[:, :, lo:hi] for gv in work.value]
    ck, cs, ce, usage = timed('get_all_sliced', lambda: work.get_all_sliced(lo, hi), acc)
    top = timed('topk(usage, 128)', lambda: torch.topk(usage, k=128, dim=-1, sorted=True)[1].flatten(), acc)
    pk = timed('gather prototype keys', lambda: (ck[:, :, top], ce[:, :, top]), acc)
    sim = timed('get_similarity 8100 x 128', lambda: get_similarity(ck, cs, pk[0], pk[1]), acc)
    aff = timed('do_softmax (dense)', lambda: do_softmax(sim), acc)
    pv = timed('readout values (2560 x 8100 @ 8100 x 128)', lambda: m._readout(aff, cand_v[0]), acc)
    ps = timed('readout shrinkage', lambda: m._readout(aff, cs), acc)
    timed('sieve_by_range', lambda: work.sieve_by_range(lo, hi, min_size=m.min_work_elements + hw), acc)
    timed('long_mem.add', lambda: m.long_mem.add(pk[0], [pv], ps, selection=None, objects=None), acc)
for k_, v_ in acc.items():
    print(f'{k_:48s} {sorted(v_)[len(v_) // 2]:9.1f} us')
print('total', round(sum(sorted(v_)[len(v_) // 2] for v_ in acc.values()), 1), 'us')
