"""What the memory-maintenance calls cost on the GPU box (they are NOT on the per-frame path: add_memory runs every
mem_every = 5th frame, consolidation when working memory is full, eviction when long-term memory is full --
tracker/inference/memory_manager.py:152-190,211-286): wall-clock per call with a device synchronise on both sides, XMem's
default bounds (config.yaml: max_mid 10, min_mid 5, 128 prototypes, max_long 10000) at the DAVIS shape, 5 objects."""
import sys, os, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vos_e_sam_b200 as vos
from tests import synth

dev = torch.device('cuda')
h, w, n_obj, cv = 30, 54, 5, 512
max_long = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
cfg = dict(hidden_dim=64, top_k=30, enable_long_term=True, enable_long_term_count_usage=True, max_mid_term_frames=10,
           min_mid_term_frames=5, num_prototypes=128, max_long_term_elements=max_long)
m = vos.MemoryManager(cfg)
g = torch.Generator().manual_seed(3)
plain, consolidate, evict, match = [], [], [], []
for step in range(400):
    k, s, e = synth.keys(g, h * w)
    v = torch.randn(1, n_obj, cv, h, w, generator=g)
    args = (k.view(1, 64, h, w).to(dev), s.view(1, 1, h, w).to(dev), v.to(dev), list(range(1, n_obj + 1)))
    sel = e.view(1, 64, h, w).to(dev)
    before = (m.work_mem.size, m.long_mem.size if m.long_mem.engaged() else 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.add_memory(*args, selection=sel)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e6
    after = (m.work_mem.size, m.long_mem.size)
    if after[0] < before[0] + h * w:
        (evict if after[1] < before[1] + 128 else consolidate).append(dt)
    else:
        plain.append(dt)
    qk, qe = synth.query(g, h, w)
    qk, qe = qk.to(dev), qe.to(dev)
    for _ in range(2):      # (usage for the next consolidation)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.match_memory(qk, qe)
        torch.cuda.synchronize()
        match.append((time.perf_counter() - t0) * 1e6)
med = lambda x: round(statistics.median(x), 1) if x else None
print(f'max_long={max_long}: add_memory plain {med(plain)} us ({len(plain)} calls) | with consolidation {med(consolidate)} us '
      f'({len(consolidate)}) | with eviction + consolidation {med(evict)} us ({len(evict)}) | match_memory (eager call, synchronised) '
      f'{med(match)} us; final sizes work {m.work_mem.size} long {m.long_mem.size}')
