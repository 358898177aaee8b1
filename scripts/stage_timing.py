"""Per-kernel CUDA-event times of one match_memory on the davis5 workload (GPU box)."""
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
for do_flush in (True, False):
    rec = []
    for it in range(30):
        qk, qe = qs[it % 4]
        e = [ev() for _ in range(5)]
        if do_flush: flush.fill_(it & 0xff)
        # events must exist (be recorded once) before their handle is valid
        for x in e: x.record()
        torch.cuda.synchronize()
        N.lib.vosmem_debug_set_stage_events(e[0].cuda_event, e[1].cuda_event, e[2].cuda_event, e[3].cuda_event)
        if do_flush: flush.fill_(it & 0xff)
        ops.match(qk.flatten(2)[0], qe.flatten(2)[0], seg, vals, 2560, 30, out=out)
        e[4].record()
        torch.cuda.synchronize()
        rec.append([e[i].elapsed_time(e[i + 1]) * 1e3 for i in range(4)])
    N.lib.vosmem_debug_set_stage_events(None, None, None, None)
    rec = rec[5:]
    names = ['pack_query', 'select_tc', 'merge+readout', 'tail']
    print('L2 flushed' if do_flush else 'L2 warm', {n: round(statistics.median(r[i] for r in rec), 1) for i, n in enumerate(names)},
          'total', round(statistics.median(sum(r) for r in rec), 1), 'us')
