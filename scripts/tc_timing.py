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
names = ['prod_wait_empty', 'mma_wait_tempty', 'mma_wait_full', 'mma_issue', 'mma_total', 'e0_wait', 'e0_relieve', 'e1_wait', 'e1_relieve', 'e2_wait', 'e2_relieve', 'e3_wait', 'e3_relieve', 'e0_loop', 'e0_total', 'first_wait', 'e0_ld(incl wait)', 'e0_groupmax', 'e0_append(incl relieve)', 'e0_active_groups', 'e0_relieve_calls', 'prologue', 'cta_total', 'g_entry', 'g_exit', 'x0_vote_wait', 'x1_group_loop', 'e0_tiles']
print('CTAs', d.shape[0])
st = d[:, 15]
print('warp1 first-tile wait cycles: mean', float(st.float().mean()), 'max', int(st.max()))
for i, nm in enumerate(names):
    col = d[:, i].float()
    print(f'{nm:16s} mean {col.mean():10.0f}  min {col.min():10.0f}  max {col.max():10.0f}')

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
