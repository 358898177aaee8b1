"""Per-role cycle counters of select_tc_kernel on the davis5 shape (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vos_e_sam_b200 as vos
from vos_e_sam_b200 import ops, _native as N
from tests import synth
g = torch.Generator().manual_seed(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16200
h = int(sys.argv[2]) if len(sys.argv) > 2 else 30
w = int(sys.argv[3]) if len(sys.argv) > 3 else 54
k, s, _ = synth.keys(g, n)
store = vos.KeyValueMemoryStore(False)
store.add(k.cuda(), [], s.cuda(), None, None)
qk, qe = synth.query(g, h, w)
q2, e2 = qk.cuda().flatten(2)[0], qe.cuda().flatten(2)[0]
seg = [store.key_segment(0, n)]
for _ in range(3): ops.select_topk(q2, e2, seg, 30, path=N.PATH_TCGEN05)
dbg = torch.zeros(4096 * 32, dtype=torch.int64, device='cuda')
N.lib.vosmem_debug_set_timing_buffer(dbg.data_ptr())
ops.select_topk(q2, e2, seg, 30, path=N.PATH_TCGEN05)
torch.cuda.synchronize()
N.lib.vosmem_debug_set_timing_buffer(None)
d = dbg.view(-1, 32).cpu()
d = d[(d != 0).any(1)]
names = {0: 'prod_wait_empty', 1: 'mma_wait_tempty', 2: 'mma_wait_full', 3: 'mma_issue', 4: 'mma_total',
         5: 'scan0_wait_tfull', 7: 'scan1_wait_tfull', 9: 'scan2_wait_tfull', 11: 'scan3_wait_tfull',
         6: 'app0_relieve', 8: 'app1_relieve', 10: 'app2_relieve', 12: 'app3_relieve',
         13: 'scan0_total', 16: 'scan0_scan', 17: 'scan0_post(wait ring)', 14: 'app0_total(incl handoff)', 26: 'app0_loop',
         25: 'app0_wait_post', 18: 'app0_append(incl relieve)', 19: 'app0_active_groups', 20: 'app0_relieve_calls',
         15: 'first_wait', 21: 'prologue', 22: 'cta_total'}
print('CTAs', d.shape[0])
st = d[:, 15]
print('warp1 first-tile wait cycles: mean', float(st.float().mean()), 'max', int(st.max()))
for i, nm in sorted(names.items()):
    col = d[:, i].float()
    print(f'{nm:28s} mean {col.mean():10.0f}  min {col.min():10.0f}  max {col.max():10.0f}')

start, end = d[:, 23].double(), d[:, 24].double()
t0 = float(start.min())
print(f'global timer (us): first CTA start 0, last CTA start {float(start.max()) - t0:.0f} ns, first CTA end {float(end.min()) - t0:.0f} ns, '
      f'last CTA end {float(end.max()) - t0:.0f} ns')
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(5):
    ev0.record(); ops.select_topk(q2, e2, seg, 30, path=N.PATH_TCGEN05); ev1.record(); torch.cuda.synchronize()
    ts.append(ev0.elapsed_time(ev1) * 1e3)
print('select_topk (pack + select + merge) event time us:', ts)
