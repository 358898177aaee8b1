"""Frames in flight: between two add_memory calls (mem_every - 1 of mem_every frames) the match_memory calls of
consecutive frames are independent of each other (the query comes from the image encoder, the memory does not change), so
their kernels may overlap.  This times CUDA-graph replays of one frame's kernels (DAVIS shape, 5 objects) issued
round-robin on 1, 2 and 3 streams -- each stream has its own workspace -- WITHOUT an L2 flush between frames (a shared
flush would serialise the streams), so compare the lines with each other, not with bench.py's flushed `value`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import vos_e_sam_b200 as vos
from vos_e_sam_b200 import ops
from tests import synth

dev = torch.device('cuda:0')
shape = bench.SHAPES['davis5']
h, w, n_obj = shape['h'], shape['w'], shape['n_obj']
g = torch.Generator().manual_seed(5)
mgr = vos.MemoryManager(bench.xmem_config(vosmem_value_dtype='bf16'))
bench.fill_memory(mgr, g, h, w, shape['frames'], n_obj, dev)
FRAMES = 600

for n_streams in (1, 2, 3):
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    graphs, keep = [], []
    for s in streams:
        qk, qe = (t.to(dev) for t in synth.query(g, h, w))
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            (p,), _ = mgr._plan_match(qk, qe)
            run = lambda p=p: ops.match(p.qk, p.qe, p.segments, p.values, p.rows, bench.TOP_K, out=p.out)
            run()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                run()
        graphs.append(gr)
        keep.append(p)
    torch.cuda.synchronize()

    def burst(frames):
        start = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in streams]
        start.record()
        for s in streams:
            s.wait_event(start)
        for i in range(frames):
            with torch.cuda.stream(streams[i % n_streams]):
                graphs[i % n_streams].replay()
        for s, e in zip(streams, ends):
            with torch.cuda.stream(s):
                e.record()
        torch.cuda.synchronize()
        return max(start.elapsed_time(e) for e in ends)
    burst(60)
    ms = min(burst(FRAMES) for _ in range(3))
    print(f'{n_streams} stream(s): {ms / FRAMES * 1e3:7.2f} us per frame, {FRAMES / ms * 1e3:9.0f} query-frames/s '
          f'(no L2 flush, {FRAMES} frames)')
