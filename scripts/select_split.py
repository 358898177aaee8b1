"""select_tc kernel vs split-merge kernel time at a given shape (GPU box): python scripts/select_split.py N H W"""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vos_e_sam_b200 as vos
from vos_e_sam_b200 import ops, _native as N
from tests import synth
n, h, w = (int(a) for a in sys.argv[1:4])
g = torch.Generator().manual_seed(1)
k, s, _ = synth.keys(g, n)
store = vos.KeyValueMemoryStore(False)
store.add(k.cuda(), [], s.cuda(), None, None)
seg = [store.key_segment(0, n)]
qs = [tuple(x.cuda().flatten(2)[0] for x in synth.query(g, h, w)) for _ in range(4)]
flush = torch.empty(512 * 2 ** 20, dtype=torch.uint8, device='cuda')
ev = lambda: torch.cuda.Event(enable_timing=True)
rec = {'pack': [], 'select': [], 'merge_splits': []}
for it in range(30):
    qk, qe = qs[it % 4]
    e = [ev() for _ in range(4)]
    for x in e: x.record()
    torch.cuda.synchronize()
    N.lib.vosmem_debug_set_stage_events(e[0].cuda_event, e[1].cuda_event, e[2].cuda_event, e[3].cuda_event)
    flush.fill_(it & 0xff)
    ops.select_topk(qk, qe, seg, 30)
    torch.cuda.synchronize()
    N.lib.vosmem_debug_set_stage_events(None, None, None, None)
    rec['pack'].append(e[0].elapsed_time(e[1]) * 1e3)
    rec['select'].append(e[1].elapsed_time(e[2]) * 1e3)
    rec['merge_splits'].append(e[2].elapsed_time(e[3]) * 1e3)
print(n, h, w, {k: round(statistics.median(v[5:]), 1) for k, v in rec.items()}, 'us (median)')
