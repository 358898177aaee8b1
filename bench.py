#!/usr/bin/env python
"""Benchmark of the space-time memory readout (MemoryManager.match_memory) -- contract in the task statement.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload davis5|lvos_sharded]

One JSON line on stdout (rank 0).  A "step" is one query frame: one match_memory call of one sequence.

  value   query-frames/s with memory AND query resident in HBM; per-step CUDA-event times on the launch stream,
          L2 flushed (512 MiB write) before every timed step; max over ranks; whole-job aggregate over N GPUs.
  e2e     the same through the public API (MemoryManager.match_memory) with HOST buffers: pinned H2D of the
          query key / selection and D2H of the readout of EVERY frame inside the timed region; two frames in
          flight (the D2H copy of frame i overlaps the kernels of frame i + 1), wall clock over K frames.
  roofline      the dominant kernel (softmax_readout_kernel, HBM-bound) from its own per-launch event times
  roofline_similarity   the fused tcgen05 similarity + selection stage against the measured bf16 peak
  cpu_baseline  the oracle port of the reference (torch CPU) on this box's host cores, bounded sample

N > 1 (launched by torchrun, one rank per GPU): independent sequences, one per rank, no data-path collective
 -> "scaling": "weak".  --workload lvos_sharded instead shards one LVOS-scale long-term
bank along N with an NCCL all-gather of the (score, index) candidates (section 8e-2).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CK, CV, TOP_K = 64, 512, 30
DAVIS = dict(h=30, w=54, frames=10, n_obj=5)                       # BASELINE.json configs[1]
LONGV = dict(h=30, w=54, frames=9, n_obj=1, n_long=10_000)          # BASELINE.json configs[2]: long-term memory engaged
BATCH = int(os.environ.get('VOSMEM_BENCH_BATCH', 11))                                                         # sequences per GPU of --workload davis_batch (configs[4]): 11 x 13 query tiles = 143 CTAs
LVOS = dict(h=68, w=120, n_long=100_000, work_frames=0, n_obj=1)   # BASELINE.json configs[3]
METRIC = 'memory_readout_query_frames_per_sec'
UNIT = 'query-frames/s'


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], tflops=p['bf16_tflops'], source='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, tflops=1590.0, source='fallback (B200_PROFILING.md)')


def ncu_traffic(workload, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/r1_traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, 'profiles', 'r1_traffic.json')))[workload].get(kernel)
    except Exception:
        return None


def xmem_config(**over):
    cfg = dict(hidden_dim=64, top_k=TOP_K, enable_long_term=True, enable_long_term_count_usage=True,
               max_mid_term_frames=10 ** 6, min_mid_term_frames=5, num_prototypes=128,
               max_long_term_elements=10 ** 7)
    cfg.update(over)
    return cfg


def fill_memory(mgr, gen, h, w, frames, n_obj, device, n_long=0):
    """Synthetic memory of the DAVIS shape: `frames` memory frames of h*w tokens, n_obj objects in one group."""
    from tests import synth
    for _ in range(frames):
        k, s, e = synth.keys(gen, h * w)
        v = torch.randn(1, n_obj, CV, h, w, generator=gen)
        mgr.add_memory(k.view(1, CK, h, w).to(device), s.view(1, 1, h, w).to(device), v.to(device),
                       list(range(1, n_obj + 1)), selection=e.view(1, CK, h, w).to(device))
    if n_long:
        k, s, _ = synth.keys(gen, n_long)
        v = torch.randn(n_obj, CV, n_long, generator=gen)
        if device == 'cpu':
            mgr.long_mem.append(k, [v], s, None, None)
        else:
            mgr.long_mem.add(k.to(device), [v.to(device)], s.to(device), None, None)


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region: NVML every 5 ms when pynvml is importable,
    else the nvidia-smi query of the B200_PROFILING.md recipe (one sample per call, ~100 ms each)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index):
        self.index, self.rows, self.stop, self.thread = index, [], threading.Event(), None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(visible.split(',')[index]) if visible and visible.split(',')[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.source = 'nvml'
        except Exception:
            self.source = 'nvidia-smi'

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, 'nvmlDeviceGetCurrentClocksEventReasons') \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = [getattr(n, 'nvmlClocksThrottleReasonHwSlowdown', 0x8), getattr(n, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                getattr(n, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20), getattr(n, 'nvmlClocksThrottleReasonSwPowerCap', 0x4)]
        self.rows.append([str(sm), str(mx)] + ['Active' if r & b else 'Not Active' for b in bits])

    def _sample_smi(self):
        out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                              '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
        self.rows.append([c.strip() for c in out.strip().split(',')])

    def _run(self):
        while not self.stop.is_set():
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:
                pass
            self.stop.wait(0.005 if self.nvml else 0.1)

    def __enter__(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=10)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for nm, val in zip(self.NAMES, r[2:6]):
                    if val.lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                continue
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx or None, reasons=sorted(reasons),
                    samples=len(sm), source=self.source)


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(workload, seconds_budget=12.0, min_calls=3, max_calls=40):
    """The reference's own algorithm (oracle port, torch CPU, all host threads) on the same workload.
    Returns (query-frames/s, calls, threads)."""
    from oracle import readout_oracle as orc
    from tests import synth
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234 + 2)
    ref = orc.Readout(xmem_config())
    if workload in ('davis5', 'davis_batch'):
        fill_memory(ref, g, DAVIS['h'], DAVIS['w'], DAVIS['frames'], DAVIS['n_obj'], 'cpu')
        h, w = DAVIS['h'], DAVIS['w']
    elif workload == 'long_video':
        fill_memory(ref, g, LONGV['h'], LONGV['w'], LONGV['frames'], LONGV['n_obj'], 'cpu', n_long=LONGV['n_long'])
        h, w = LONGV['h'], LONGV['w']
    else:
        # bounded sample of the LVOS workload: a quarter of the query rows against the full bank
        fill_memory(ref, g, 17, 120, 1, 1, 'cpu', n_long=LVOS['n_long'])
        h, w = 17, 120
    qk, qe = synth.query(g, h, w)
    ref.match_memory(qk, qe)  # warm-up (thread pool, allocator)
    times = []
    t_end = time.perf_counter() + seconds_budget
    while len(times) < max_calls and (len(times) < min_calls or time.perf_counter() < t_end):
        t0 = time.perf_counter()
        ref.match_memory(qk, qe)
        times.append(time.perf_counter() - t0)
    scale = 1.0 if workload != 'lvos_sharded' else (17 * 120) / (LVOS['h'] * LVOS['w'])
    return scale / statistics.median(times), len(times), torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    rate, calls, threads = cpu_reference_rate(args.workload, seconds_budget=max(10.0, 0.5 * args.steps))
    sample = (f'{calls} match_memory calls of the oracle port (torch CPU fp32) on the {args.workload} workload, '
              f'median call time' + ('' if args.workload != 'lvos_sharded' else '; 1/4 of the query rows, rate scaled'))
    line = dict(metric=METRIC, value=rate, unit=UNIT, n_gpus=args.gpus, steps=calls, warmup=1,
                ms_per_step=1000.0 / rate, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic', impl='reference', config=workload_config(args.workload, args.gpus),
                cpu_baseline=dict(value=rate, unit=UNIT, cores=threads, kind='port', sample=sample),
                e2e=dict(value=rate, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def workload_config(workload, n_gpus):
    if workload == 'davis_batch':
        c = workload_config('davis5', n_gpus)
        c.update(workload='davis2017_multiobject_readout_batched', sequences_per_gpu=BATCH,
                 launch='one CUDA-graph replay per step; a step = one frame of each of the sequences (vosmem_match_batch)',
                 parallelism=f'dp{n_gpus} x {BATCH} independent sequences per GPU')
        return c
    if workload == 'long_video':
        return dict(workload='long_video_longterm_readout', hw=LONGV['h'] * LONGV['w'],
                    memory_elements=LONGV['n_long'] + LONGV['frames'] * LONGV['h'] * LONGV['w'], long_term_elements=LONGV['n_long'],
                    objects=1, ck=CK, cv=CV, top_k=TOP_K, value_storage='bf16',
                    similarity='bf16 hi/lo split x3 -> fp32 (tcgen05)', l2='flushed before every timed step',
                    launch='one CUDA-graph replay per step', parallelism=f'dp{n_gpus} (one independent sequence per GPU)')
    if workload == 'davis5':
        return dict(workload='davis2017_multiobject_readout', hw=DAVIS['h'] * DAVIS['w'],
                    memory_elements=DAVIS['frames'] * DAVIS['h'] * DAVIS['w'], objects=DAVIS['n_obj'], ck=CK, cv=CV,
                    top_k=TOP_K, sensory_hidden='1x5x64x30x54 held by the manager', value_storage='bf16',
                    similarity='bf16 hi/lo split x3 -> fp32 (tcgen05)', l2='flushed before every timed step', launch='one CUDA-graph replay per step',
                    parallelism=f'dp{n_gpus} (one independent sequence per GPU)')
    return dict(workload='lvos1080p_longterm_sharded_readout', hw=LVOS['h'] * LVOS['w'], memory_elements=LVOS['n_long'],
                objects=1, ck=CK, cv=CV, top_k=TOP_K, value_storage='bf16', l2='flushed before every timed step',
                parallelism=f'long-term bank sharded along N over {n_gpus} GPU(s), candidate exchange: see config.exchange')


# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (there is no CPU path in vos_e_sam_b200)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import vos_e_sam_b200 as vos
    from vos_e_sam_b200 import ops
    from tests import synth

    K, W = args.steps, args.warmup
    g = torch.Generator().manual_seed(1234 + 2 + rank)
    sharded = args.workload == 'lvos_sharded'
    n_seq = 1
    if sharded:
        from vos_e_sam_b200.sharded import ShardedLongTermReadout
        h, w, n_obj = LVOS['h'], LVOS['w'], 1
        engine = ShardedLongTermReadout(xmem_config(vosmem_exchange=args.exchange, vosmem_shard=args.shard), rank, world, dev)
        gl = torch.Generator().manual_seed(1234 + 4)     # same bank on every rank; each keeps its shard
        k, s, _ = synth.keys(gl, LVOS['n_long'])
        v = torch.randn(n_obj, CV, LVOS['n_long'], generator=gl)
        engine.load_long_term(k, s, v)
        n_mem = LVOS['n_long']
    else:
        shape = LONGV if args.workload == 'long_video' else DAVIS
        h, w, n_obj = shape['h'], shape['w'], shape['n_obj']
        n_seq = BATCH if args.workload == 'davis_batch' else 1
        mgrs = []
        for _ in range(n_seq):
            m = vos.MemoryManager(xmem_config(vosmem_value_dtype='bf16'))
            fill_memory(m, g, h, w, shape['frames'], n_obj, dev, n_long=shape.get('n_long', 0))
            m.create_hidden_state(n_obj, torch.empty(1, CK, h, w, device=dev))
            mgrs.append(m)
        mgr = mgrs[0]
        n_mem = mgr.work_mem.size + (mgr.long_mem.size if mgr.long_mem.engaged() else 0)
    hw = h * w
    rows = n_obj * CV

    # a small pool of distinct query frames, host (pinned) and device copies
    pool = 4
    # pool entry = the query frames of one step: n_seq x (key, selection), one pinned host tensor each
    host_q = [tuple(torch.stack(x).pin_memory() for x in zip(*[synth.query(g, h, w) for _ in range(n_seq)]))
              for _ in range(pool)]                                   # each n_seq x 1 x CK x h x w
    if n_seq == 1:
        host_q = [(a[0], b[0]) for a, b in host_q]
    dev_q = [(a.to(dev), b.to(dev)) for a, b in host_q]
    flush = torch.empty(512 * 2 ** 20, dtype=torch.uint8, device=dev)
    host_out = torch.empty((n_seq, n_obj, CV, h, w) if n_seq > 1 else (n_obj, CV, h, w), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps: stage-split so each stage has its own events ---------------------
    if not sharded:
        def single_problem(j):
            """the (single) object group of the sequence as the arguments of one vosmem_match call"""
            qk, qe = dev_q[j]
            (p,), _ = mgr._plan_match(qk, qe)
            return p

        from vos_e_sam_b200 import _native as N

        def batch_problems(j):
            """the object groups of one frame of every sequence, as one vosmem_match_batch problem list"""
            qk, qe = dev_q[j]
            return [p for m, k, e in zip(mgrs, qk, qe) for p in m._plan_match(k, e)[0]]

        def step(i, ev):
            qk, qe = dev_q[i % pool]
            flush.fill_(i & 0xFF)
            # the C call records ev[0..3] on the stream around its pack / select / readout kernels
            N.lib.vosmem_debug_set_stage_events(ev[0].cuda_event, ev[1].cuda_event, ev[2].cuda_event, ev[3].cuda_event)
            if n_seq == 1:
                p = single_problem(i % pool)
                ops.match(p.qk, p.qe, p.segments, p.values, p.rows, TOP_K, out=p.out)  # tcgen05 select (packs the query), merge+softmax+readout(+age)
            else:
                ops.match_batch(batch_problems(i % pool), TOP_K)
            ev[4].record()
        launches_per_step = 2                                   # select_tc (packs its query tiles), softmax_readout (ages life_count)
    else:
        def step(i, ev):
            qk, qe = dev_q[i % pool]
            flush.fill_(i & 0xFF)
            ev[0].record()
            engine.match(qk, qe, events=ev[1:4])    # after local select / exchange + merge / readout
            ev[4].record()
        launches_per_step = engine.launches_per_match

    def new_events():
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        for e in evs:
            e.record()          # a torch event only owns a CUDA handle once it has been recorded
        return evs

    # ---- (1) headline: each step = one CUDA-graph replay of the step's kernels (no host launch gaps) ----------------
    graphs = None
    if not sharded:
        N.lib.vosmem_debug_set_stage_events(None, None, None, None)
        graphs, graph_outputs = [], []
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for j in range(pool):
                if n_seq == 1:
                    p = single_problem(j)
                    graph_outputs.append(p)
                    run = lambda: ops.match(p.qk, p.qe, p.segments, p.values, p.rows, TOP_K, out=p.out)
                else:
                    probs = batch_problems(j)
                    graph_outputs.append(probs)      # the captured kernels write these tensors on every replay
                    run = lambda: ops.match_batch(probs, TOP_K)
                run()                                                    # warm the workspace cache outside capture
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=side):
                    run()
                graphs.append(gr)
        torch.cuda.current_stream().wait_stream(side)

        def run_step(i, ev):
            flush.fill_(i & 0xFF)
            ev[0].record()
            graphs[i % pool].replay()
            ev[4].record()
    else:
        run_step = step

    for i in range(W):
        run_step(i, new_events())
    barrier()
    events = [new_events() for _ in range(K)]
    stage_events = [new_events() for _ in range(K)]
    with ClockSampler(local_rank) as clocks:
        barrier()
        for i in range(K):
            run_step(i, events[i])
        barrier()
        # ---- (2) the same steps launched kernel by kernel with events between the kernels: per-kernel durations ----
        if not sharded:
            for i in range(W):
                step(i, new_events())
            for i in range(K):
                step(i, stage_events[i])
            barrier()
            N.lib.vosmem_debug_set_stage_events(None, None, None, None)   # the e2e calls below must not re-record them
        else:
            stage_events = events
        # ---- end-to-end through the public API with host buffers -------------------------------
        # Two frames in flight: the D2H copy of frame i (copy stream, its own pinned buffer) overlaps the H2D copy and
        # the kernels of frame i + 1; the host waits for frame i - 2's copy before that buffer is reused.
        e2e_steps = K
        copy_stream = torch.cuda.Stream()
        host_outs = [host_out, torch.empty_like(host_out).pin_memory()]
        landed = [None, None]

        def e2e_step(i):
            a, b = host_q[i % pool]
            qk_d, qe_d = a.to(dev, non_blocking=True), b.to(dev, non_blocking=True)
            if sharded:
                r = engine.match(qk_d, qe_d)
            elif n_seq == 1:
                r = mgr.match_memory(qk_d, qe_d)
            else:
                r = torch.stack(vos.match_memory_batch(mgrs, qk_d, qe_d))
            ready = torch.cuda.Event()
            ready.record()
            slot = i % 2
            if landed[slot] is not None:
                landed[slot].synchronize()      # the host has frame i - 2's readout
            copy_stream.wait_event(ready)
            with torch.cuda.stream(copy_stream):
                dst = host_outs[slot].view(rows, hw)[:r.shape[0], :r.shape[1]] if sharded else host_outs[slot]
                dst.copy_(r, non_blocking=True)
                r.record_stream(copy_stream)
                landed[slot] = torch.cuda.Event()
                landed[slot].record()

        def e2e_drain():
            for e in landed:
                if e is not None:
                    e.synchronize()
            torch.cuda.synchronize()
        for i in range(W):
            e2e_step(i)
        e2e_drain()
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step(i)
        e2e_drain()
        barrier()
        e2e_s = time.perf_counter() - t0
    step_ms = [e[0].elapsed_time(e[4]) for e in events]
    se = stage_events
    # davis5: [begin, after pack, after select, after readout, end]; sharded: [begin, after pack + select, after
    # exchange + merge, after readout, end]
    pack_ms = [e[0].elapsed_time(e[1]) for e in se] if not sharded else [0.0 for e in se]
    sel_ms = [e[1].elapsed_time(e[2]) for e in se] if not sharded else [e[0].elapsed_time(e[1]) for e in se]
    rd_ms = [e[2].elapsed_time(e[3]) for e in se]
    xchg_ms = [e[1].elapsed_time(e[2]) for e in se] if sharded else None

    total_ms = sum(step_ms)

    stats = torch.tensor([total_ms, e2e_s * 1000.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = stats.tolist()
    frames = K if sharded else K * world * n_seq       # sharded: all ranks cooperate on the same frames
    value = frames / (total_ms / 1000.0)
    e2e_value = (e2e_steps if sharded else e2e_steps * world * n_seq) / (e2e_ms / 1000.0)

    if rank != 0:
        return
    pk = peaks()
    val_bytes = 2
    n_touch = min(n_mem, hw * TOP_K)
    rd_bytes = n_seq * (rows * n_touch * val_bytes + rows * hw * 4 + hw * TOP_K * 12)
    rd_t = statistics.mean(rd_ms) / 1000.0
    sel_flops = 4.0 * n_mem * hw * CK * n_seq
    sel_t = statistics.mean(sel_ms) / 1000.0
    if sharded:
        rd_bytes = rows * min(n_mem, hw * TOP_K) * val_bytes + rows * hw * 4 + hw * TOP_K * 12   # readout is replicated
        sel_flops /= world
    roof_rd = dict(kernel='softmax_readout_kernel (merge + softmax + usage + sparse readout)', bound='hbm', achieved=rd_bytes / rd_t / 1e9, peak=pk['hbm'], unit='GB/s',
                   frac=rd_bytes / rd_t / 1e9 / pk['hbm'], traffic=ncu_traffic(args.workload, 'softmax_readout_kernel') if world == 1 and n_seq == 1 else None,
                   us_per_launch=rd_t * 1e6,
                   algorithmic_bytes=rd_bytes, peak_source=pk['source'])
    roof_sel = dict(kernel='select_tc_kernel', bound='tensor',
                    achieved=sel_flops / sel_t / 1e12, peak=pk['tflops'], unit='TFLOP/s',
                    frac=sel_flops / sel_t / 1e12 / pk['tflops'], traffic=ncu_traffic(args.workload, 'select_tc_kernel') if world == 1 and n_seq == 1 else None,
                    us_per_launch=sel_t * 1e6,
                    algorithmic_flops=sel_flops, executed_flop_multiplier=25.0 / 8.0, executed_frac=25.0 / 8.0 * sel_flops / sel_t / 1e12 / pk['tflops'], peak_source=pk['source'] + ', burst')
    dominant, other = (roof_rd, roof_sel) if rd_t >= sel_t else (roof_sel, roof_rd)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, calls, threads = cpu_reference_rate(args.workload)
        cpu = dict(value=rate, unit=UNIT, cores=threads, kind='port',
                   sample=f'{calls} match_memory calls of the oracle port (torch CPU fp32) on the same workload'
                          + ('' if not sharded else ', 1/4 of the query rows, rate scaled'))
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W,
                ms_per_step=total_ms / K, higher_is_better=True, scaling='strong' if sharded else 'weak',
                vs_baseline=None, dtype='bf16', data='synthetic',
                config=dict(workload_config(args.workload, world), **({'exchange': args.exchange, 'shard': args.shard} if sharded else {})),
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=n_seq * 2 * CK * hw * 4, d2h_bytes_per_step=n_seq * rows * hw * 4),
                gpu_launches=K * launches_per_step, clocks=clocks.summary(), roofline=dominant,
                roofline_other=other, cpu_baseline=cpu,
                stage_us=dict(pack_query=statistics.mean(pack_ms) * 1e3, select=statistics.mean(sel_ms) * 1e3,
                              readout=statistics.mean(rd_ms) * 1e3,
                              **(dict(exchange_and_merge=statistics.mean(xchg_ms) * 1e3,
                                      note='select includes the query packing kernel') if sharded else {}),
                              step_median=statistics.median(step_ms) * 1e3,
                              step_kernel_by_kernel=statistics.median(e[0].elapsed_time(e[4]) for e in se) * 1e3))
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='davis5', choices=['davis5', 'davis_batch', 'long_video', 'lvos_sharded'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--shard', default='n', choices=['n', 'queries'],
                    help='lvos_sharded: shard the key axis (north_star) or, as a control, the query rows')
    ap.add_argument('--exchange', default='nccl', choices=['nccl', 'peer'],
                    help='lvos_sharded: candidate exchange by NCCL all-gather or by peer-memory loads inside the merge kernel')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # not under torchrun: re-launch ourselves one rank per GPU
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
               '--master-addr', '127.0.0.1', '--master-port', str(29500 + os.getpid() % 1000), os.path.abspath(__file__),
               '--gpus', str(args.gpus), '--steps', str(args.steps), '--warmup', str(args.warmup),
               '--workload', args.workload, '--exchange', args.exchange, '--shard', args.shard] + (['--no-cpu-baseline'] if args.no_cpu_baseline else [])
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
