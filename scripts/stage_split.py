"""Where the time of the davis5 readout goes (GPU box): select | split merge | staged readout (softmax + gather only)
against the fused merge + readout kernel, per-kernel CUDA-event times, L2 flushed before every call."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import vos_e_sam_b200 as vos
from vos_e_sam_b200 import ops, _native as N
from tests import synth
dev = torch.device('cuda')
g = torch.Generator().manual_seed(3)
mgr = vos.MemoryManager(bench.xmem_config(vosmem_value_dtype='bf16'))
bench.fill_memory(mgr, g, 30, 54, 10, 5, dev)
work = mgr.work_mem
seg = [work.key_segment(0, work.size)]
vals = [work.value_segment(0, 0, with_usage=True)]
out = torch.empty((2560, 1620), device=dev)
qs = [tuple(x.to(dev) for x in synth.query(g, 30, 54)) for _ in range(4)]
flush = torch.empty(512 * 2 ** 20, dtype=torch.uint8, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
rec = {'select': [], 'merge_splits': [], 'staged_readout': [], 'fused_readout': []}
for it in range(40):
    qk, qe = (t.flatten(2)[0] for t in qs[it % 4])
    e = [ev() for _ in range(4)]
    for x in e: x.record()
    torch.cuda.synchronize()
    N.lib.vosmem_debug_set_stage_events(e[0].cuda_event, e[1].cuda_event, e[2].cuda_event, e[3].cuda_event)
    flush.fill_(it & 0xff)
    s, i = ops.select_topk(qk, qe, seg, 30)
    torch.cuda.synchronize()
    N.lib.vosmem_debug_set_stage_events(None, None, None, None)
    rec['select'].append(e[1].elapsed_time(e[2]) * 1e3)
    rec['merge_splits'].append(e[2].elapsed_time(e[3]) * 1e3)
    a, b = ev(), ev()
    flush.fill_(it & 0xff)
    a.record(); ops.softmax_readout(s, i, vals, 2560, out=out); b.record(); torch.cuda.synchronize()
    rec['staged_readout'].append(a.elapsed_time(b) * 1e3)
    e = [ev() for _ in range(4)]
    for x in e: x.record()
    torch.cuda.synchronize()
    N.lib.vosmem_debug_set_stage_events(e[0].cuda_event, e[1].cuda_event, e[2].cuda_event, e[3].cuda_event)
    flush.fill_(it & 0xff)
    ops.match(qk, qe, seg, vals, 2560, 30, out=out)
    torch.cuda.synchronize()
    N.lib.vosmem_debug_set_stage_events(None, None, None, None)
    rec['fused_readout'].append(e[2].elapsed_time(e[3]) * 1e3)
print({k: round(statistics.median(v[5:]), 1) for k, v in rec.items()}, 'us (median)')
