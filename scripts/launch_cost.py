"""Launch + scheduling cost of the select_tc grid alone (VOSMEM_TC_EXP=4: every CTA returns at entry) against the full
kernel, as CUDA-graph replays of select_topk (select + merge kernels) at the DAVIS shape."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vos_e_sam_b200 as vos
from vos_e_sam_b200 import ops, _native as N
from tests import synth
g = torch.Generator().manual_seed(1)
n, h, w = 16200, 30, 54
k, s, _ = synth.keys(g, n)
store = vos.KeyValueMemoryStore(False)
store.add(k.cuda(), [], s.cuda(), None, None)
qk, qe = synth.query(g, h, w)
q2, e2 = qk.cuda().flatten(2)[0], qe.cuda().flatten(2)[0]
seg = [store.key_segment(0, n)]
out = (torch.empty((h * w, 30), device='cuda'), torch.empty((h * w, 30), dtype=torch.int64, device='cuda'))
side = torch.cuda.Stream()
with torch.cuda.stream(side):     # (the workspace is per stream: warm up on the stream the graph is captured on)
    for _ in range(3): ops.select_topk(q2, e2, seg, 30, path=N.PATH_TCGEN05, out=out)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr, stream=side):
    ops.select_topk(q2, e2, seg, 30, path=N.PATH_TCGEN05, out=out)
flush = torch.empty(512 * 2 ** 20, dtype=torch.uint8, device='cuda')
ts = []
for it in range(30):
    flush.fill_(it & 0xff)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
print('VOSMEM_TC_EXP =', os.environ.get('VOSMEM_TC_EXP', '0'), ': graph replay of select + merge, median us', round(statistics.median(ts[5:]), 2))
