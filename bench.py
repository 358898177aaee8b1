#!/usr/bin/env python
"""Benchmark of the space-time memory readout (MemoryManager.match_memory) -- contract in the task statement.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload davis5|long_video|davis_batch|lvos_sharded]

One JSON line on stdout (rank 0).  A "step" is one query frame: one match_memory call of one sequence.

  value   query-frames/s of the headline workload (BASELINE.json configs[1], `davis5`) with memory AND query resident in
          HBM; per-step CUDA-event times on the launch stream, L2 flushed (512 MiB write) before every timed step; one
          CUDA-graph replay per step; max over ranks; whole-job aggregate over N GPUs (independent sequences, one per
          rank, no data-path collective: SURVEY.md section 8e-1 -> "scaling": "weak").
  e2e     the same through the public API (MemoryManager.match_memory) with HOST buffers: pinned H2D of the query key /
          selection and D2H of the readout of EVERY frame inside the timed region; two frames in flight.  `pcie_ceiling`
          is what the box's PCIe allows for that D2H traffic (all ranks copying at once, no kernels).
  roofline        the dominant kernel, ALGORITHMIC bytes counted from the value rows actually touched (unique survivors)
  roofline_other  the other kernel; roofline_step: the whole step against SURVEY.md section 8d's bytes / flops
  cpu_baseline        the reference (oracle/_ref, staged copy of the unmodified modules) or its oracle port on the host cores
  gpu_eager_baseline  the reference's own torch op sequence on CUDA tensors on the same GPU (secondary line)
  key_projection      SURVEY 8f-4: the tcgen05 KeyProjection against the reference module on cuDNN, same GPU
  maintenance         add_memory / consolidation / eager-call wall times (not on the per-frame path)
  other_workloads     BASELINE.json configs[2] (`long_video`) and configs[4] (`davis_batch`), short runs of the same kind
  sharded         BASELINE.json configs[3]: ONE LVOS-scale long-term bank sharded along N over the N ranks (strong scaling),
                  with the in-run 1-GPU unsharded time next to it, efficiency = T1 / (N * TN), and the parity of the sharded
                  result against the unsharded one
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
BATCH = int(os.environ.get('VOSMEM_BENCH_BATCH', 11))               # sequences per GPU of `davis_batch` (configs[4]): 11 x 13 query tiles = 143 CTAs
LVOS = dict(h=68, w=120, n_long=100_000, n_obj=1)                   # BASELINE.json configs[3]
METRIC = 'memory_readout_query_frames_per_sec'
UNIT = 'query-frames/s'
SHAPES = dict(davis5=DAVIS, davis_batch=DAVIS, long_video=LONGV)


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], tflops=p['bf16_tflops'], tflops_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, tflops=1590.0, tflops_sustained=1400.0, source='fallback (B200_PROFILING.md)')


def ncu_traffic(workload, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/*_traffic.json, newest round)."""
    for name in ('r2_traffic.json', 'r1_traffic.json'):
        try:
            v = json.load(open(os.path.join(ROOT, 'profiles', name)))[workload].get(kernel)
            if v is not None:
                return v
        except Exception:
            continue
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
        lm = mgr.long_mem
        if hasattr(lm, 'append'):                       # oracle port
            lm.append(k, [v], s, None, None)
        else:                                           # product manager / reference manager
            lm.add(k.to(device), [v.to(device)], s.to(device), None, None)


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
# reference arm / CPU baseline
def reference_manager(workload, device='cpu'):
    """(manager, kind, h, w, scale): the reference's MemoryManager from the staged copy (oracle/_ref), else the oracle port,
    filled with the workload's memory.  lvos_sharded: a bounded sample -- 1/4 of the query rows against the full bank."""
    from oracle import build_ref
    if build_ref.available():
        Manager, _, _ = build_ref.load()
        kind = 'reference'
    else:
        from oracle import readout_oracle as orc
        Manager, kind = orc.Readout, 'port'
    g = torch.Generator().manual_seed(1234 + 2)
    ref = Manager(xmem_config())
    scale = 1.0
    if workload in SHAPES:
        sh = SHAPES[workload]
        fill_memory(ref, g, sh['h'], sh['w'], sh['frames'], sh['n_obj'], device, n_long=sh.get('n_long', 0))
        h, w = sh['h'], sh['w']
    else:
        fill_memory(ref, g, 17, 120, 1, 1, device, n_long=LVOS['n_long'])
        h, w = 17, 120
        scale = (17 * 120) / (LVOS['h'] * LVOS['w'])
    return ref, kind, h, w, scale


def cpu_reference_rate(workload, steps=None, warmup=1, seconds_budget=12.0, min_calls=3, max_calls=40):
    """The reference's own implementation of the path on the host cores (all threads).  With `steps` given, exactly that
    many timed calls after `warmup` untimed ones; else as many as fit the time budget.
    Returns (query-frames/s, calls, threads, kind, ms per call list)."""
    from tests import synth
    torch.set_num_threads(os.cpu_count() or 1)
    ref, kind, h, w, scale = reference_manager(workload)
    g = torch.Generator().manual_seed(99)
    qk, qe = synth.query(g, h, w)
    for _ in range(max(1, warmup)):
        ref.match_memory(qk, qe)                     # warm-up (thread pool, allocator)
    times = []
    t_end = time.perf_counter() + seconds_budget
    while (len(times) < steps) if steps else (len(times) < max_calls and (len(times) < min_calls or time.perf_counter() < t_end)):
        t0 = time.perf_counter()
        ref.match_memory(qk, qe)
        times.append(time.perf_counter() - t0)
    return scale / statistics.median(times), len(times), torch.get_num_threads(), kind, times


def cpu_sample_text(kind, calls, workload):
    what = 'the unmodified reference MemoryManager (oracle/_ref, torch CPU fp32)' if kind == 'reference' else \
        'the oracle port of the reference (torch CPU fp32)'
    return (f'{calls} match_memory calls of {what} on the {workload} workload, median call time'
            + ('' if workload != 'lvos_sharded' else '; 1/4 of the query rows against the full bank, rate scaled'))


def run_reference(args, rank):
    if rank != 0:
        return
    # each step = one match_memory call of the workload (a bounded sample for lvos_sharded); K timed after W warm-ups
    # (bounded so that the run ends within minutes: the default 200 steps are ~30 s of davis5 calls)
    steps = min(args.steps, 200 if args.workload != 'lvos_sharded' else 20)
    warm = min(args.warmup, 10)
    rate, calls, threads, kind, times = cpu_reference_rate(args.workload, steps=steps, warmup=warm)
    line = dict(metric=METRIC, value=rate, unit=UNIT, n_gpus=args.gpus, steps=calls, warmup=warm,
                ms_per_step=1000.0 / rate, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic', impl='reference', config=workload_config(args.workload, args.gpus),
                cpu_baseline=dict(value=rate, unit=UNIT, cores=threads, kind=kind, sample=cpu_sample_text(kind, calls, args.workload),
                                  timed_calls=calls),
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
                parallelism=f'long-term bank sharded along N over {n_gpus} GPU(s), readout sharded over the query rows')


# ------------------------------------------------------------------------------------------------
class Ctx:
    """What every measurement needs: device, ranks, the L2 flush buffer, barriers."""

    def __init__(self, rank, world, local_rank):
        import torch.distributed as dist
        self.rank, self.world, self.dist = rank, world, dist
        self.dev = torch.device('cuda', local_rank)
        self.flush = torch.empty(512 * 2 ** 20, dtype=torch.uint8, device=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = torch.tensor(vals, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()


def new_events(n=5):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
    for e in evs:
        e.record()          # a torch event only owns a CUDA handle once it has been recorded
    return evs


def measure_dp(ctx, workload, K, W, want_e2e, sample_clocks):
    """Independent sequences, one (or BATCH) per rank: graph-replayed steps, per-kernel stage times, optional e2e."""
    import vos_e_sam_b200 as vos
    from vos_e_sam_b200 import ops, _native as N
    from tests import synth
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    shape = SHAPES[workload]
    h, w, n_obj = shape['h'], shape['w'], shape['n_obj']
    n_seq = BATCH if workload == 'davis_batch' else 1
    g = torch.Generator().manual_seed(1234 + 2 + rank)
    mgrs = []
    for _ in range(n_seq):
        m = vos.MemoryManager(xmem_config(vosmem_value_dtype='bf16'))
        fill_memory(m, g, h, w, shape['frames'], n_obj, dev, n_long=shape.get('n_long', 0))
        m.create_hidden_state(n_obj, torch.empty(1, CK, h, w, device=dev))
        mgrs.append(m)
    mgr = mgrs[0]
    n_mem = mgr.work_mem.size + (mgr.long_mem.size if mgr.long_mem.engaged() else 0)
    hw, rows = h * w, n_obj * CV

    pool = 4     # distinct query frames, host (pinned) and device copies; pool entry = the query frames of one step
    host_q = [tuple(torch.stack(x).pin_memory() for x in zip(*[synth.query(g, h, w) for _ in range(n_seq)]))
              for _ in range(pool)]                                   # each n_seq x 1 x CK x h x w
    if n_seq == 1:
        host_q = [(a[0], b[0]) for a, b in host_q]
    dev_q = [(a.to(dev), b.to(dev)) for a, b in host_q]
    host_out = torch.empty((n_seq, n_obj, CV, h, w) if n_seq > 1 else (n_obj, CV, h, w), dtype=torch.float32).pin_memory()
    flush = ctx.flush

    def single_problem(j):
        """the (single) object group of the sequence as the arguments of one vosmem_match call"""
        qk, qe = dev_q[j]
        (p,), _ = mgr._plan_match(qk, qe)
        return p

    def batch_problems(j):
        """the object groups of one frame of every sequence, as one vosmem_match_batch problem list"""
        qk, qe = dev_q[j]
        return [p for m, k, e in zip(mgrs, qk, qe) for p in m._plan_match(k, e)[0]]

    def step(i, ev):
        flush.fill_(i & 0xFF)
        # the C call records ev[0..3] on the stream around its pack / select / readout kernels
        N.lib.vosmem_debug_set_stage_events(ev[0].cuda_event, ev[1].cuda_event, ev[2].cuda_event, ev[3].cuda_event)
        if n_seq == 1:
            p = single_problem(i % pool)
            ops.match(p.qk, p.qe, p.segments, p.values, p.rows, TOP_K, out=p.out)
        else:
            ops.match_batch(batch_problems(i % pool), TOP_K)
        ev[4].record()

    # ---- (1) headline: each step = one CUDA-graph replay of the step's kernels (no host launch gaps) ----
    N.lib.vosmem_debug_set_stage_events(None, None, None, None)
    graphs, graph_outputs = [], []
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for j in range(pool):
            if n_seq == 1:
                p = single_problem(j)
                graph_outputs.append(p)
                run = lambda p=p: ops.match(p.qk, p.qe, p.segments, p.values, p.rows, TOP_K, out=p.out)
            else:
                probs = batch_problems(j)
                graph_outputs.append(probs)      # the captured kernels write these tensors on every replay
                run = lambda probs=probs: ops.match_batch(probs, TOP_K)
            run()                                # allocates + initialises the workspaces outside the capture
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

    for i in range(W):
        run_step(i, new_events())
    ctx.barrier()
    events = [new_events() for _ in range(K)]
    stage_events = [new_events() for _ in range(K)]
    e2e_s = None
    with ClockSampler(dev.index) as clocks:
        ctx.barrier()
        for i in range(K):
            run_step(i, events[i])
        ctx.barrier()
        # ---- (2) the same steps launched kernel by kernel with events between the kernels: per-kernel durations ----
        for i in range(W):
            step(i, new_events())
        for i in range(K):
            step(i, stage_events[i])
        ctx.barrier()
        N.lib.vosmem_debug_set_stage_events(None, None, None, None)   # the e2e calls below must not re-record them
        # ---- (3) end to end through the public API with host buffers ----
        if want_e2e:
            # Two frames in flight: the D2H copy of frame i (copy stream, its own pinned buffer) overlaps the H2D copy and
            # the kernels of frame i + 1; the host waits for frame i - 2's copy before that buffer is reused.
            copy_stream = torch.cuda.Stream()
            host_outs = [host_out, torch.empty_like(host_out).pin_memory()]
            landed = [None, None]

            def e2e_step(i):
                a, b = host_q[i % pool]
                qk_d, qe_d = a.to(dev, non_blocking=True), b.to(dev, non_blocking=True)
                r = mgr.match_memory(qk_d, qe_d) if n_seq == 1 else torch.stack(vos.match_memory_batch(mgrs, qk_d, qe_d))
                ready = torch.cuda.Event()
                ready.record()
                slot = i % 2
                if landed[slot] is not None:
                    landed[slot].synchronize()      # the host has frame i - 2's readout
                copy_stream.wait_event(ready)
                with torch.cuda.stream(copy_stream):
                    host_outs[slot].copy_(r, non_blocking=True)
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
            ctx.barrier()
            t0 = time.perf_counter()
            for i in range(K):
                e2e_step(i)
            e2e_drain()
            ctx.barrier()
            e2e_s = time.perf_counter() - t0
    step_ms = [e[0].elapsed_time(e[4]) for e in events]
    se = stage_events
    pack_ms = [e[0].elapsed_time(e[1]) for e in se]
    sel_ms = [e[1].elapsed_time(e[2]) for e in se]
    rd_ms = [e[2].elapsed_time(e[3]) for e in se]
    total_ms, e2e_ms = ctx.max_over_ranks(sum(step_ms), (e2e_s or 0.0) * 1000.0)
    frames = K * world * n_seq

    # value rows actually touched by one frame's readout (unique survivors), counted on the device
    uniq = []
    for j in range(pool):
        for p in ([single_problem(j)] if n_seq == 1 else batch_problems(j)):
            _, idx = ops.select_topk(p.qk, p.qe, p.segments, TOP_K)
            uniq.append(int(torch.unique(idx[idx >= 0]).numel()))
    unique_rows = statistics.mean(uniq)

    return dict(workload=workload, h=h, w=w, hw=hw, rows=rows, n_obj=n_obj, n_mem=n_mem, n_seq=n_seq, K=K,
                total_ms=total_ms, value=frames / (total_ms / 1000.0), ms_per_step=total_ms / K,
                e2e_value=(frames / (e2e_ms / 1000.0)) if want_e2e else None,
                step_ms=step_ms, pack_ms=pack_ms, sel_ms=sel_ms, rd_ms=rd_ms, unique_rows=unique_rows,
                step_kernel_by_kernel_us=statistics.median(e[0].elapsed_time(e[4]) for e in se) * 1e3,
                clocks=clocks.summary() if sample_clocks else None, mgr=mgr, dev_q=dev_q)


def rooflines(r, world, pk):
    """roofline objects of a measure_dp result: readout (HBM), selection (tensor), whole step."""
    hw, rows, n_mem, n_seq, uniq = r['hw'], r['rows'], r['n_mem'], r['n_seq'], r['unique_rows']
    val_bytes = 2
    rd_t = statistics.mean(r['rd_ms']) / 1000.0
    sel_t = statistics.mean(r['sel_ms']) / 1000.0
    # readout: every touched value row once + the rows x HW output + the survivors' (weight, index)
    rd_bytes = n_seq * (rows * uniq * val_bytes + rows * hw * 4 + hw * TOP_K * 8)
    rd_bytes_8d = n_seq * (rows * min(n_mem, hw * TOP_K) * val_bytes + rows * hw * 4 + hw * TOP_K * 8)
    sel_flops = 4.0 * n_mem * hw * CK * n_seq
    single = world == 1 and n_seq == 1
    roof_rd = dict(kernel='softmax_readout_kernel (merge + softmax + usage + sparse readout)', bound='hbm',
                   achieved=rd_bytes / rd_t / 1e9, peak=pk['hbm'], unit='GB/s', frac=rd_bytes / rd_t / 1e9 / pk['hbm'],
                   traffic=ncu_traffic(r['workload'], 'softmax_readout_kernel') if single else None,
                   us_per_launch=rd_t * 1e6, algorithmic_bytes=rd_bytes,
                   algorithmic_bytes_note=f'rows x unique survivors ({uniq:.0f} value rows touched per frame, counted on the device) x 2 B '
                                          f'+ rows x HW x 4 B + HW x k x 8 B',
                   frac_section_8d_bytes=rd_bytes_8d / rd_t / 1e9 / pk['hbm'],
                   l2_to_sm_gather_bytes=n_seq * hw * TOP_K * rows * val_bytes,
                   peak_source=pk['source'])
    # What actually limits this kernel is not HBM (the touched rows are re-read ~10x out of L2) but the L2 -> SM fabric:
    # every pick moves its whole value row.  Ceiling: the LTS throughput cap of B300_MICROARCH.md (~6 300 B / cycle for the
    # whole chip) at the SM clock sampled during the run.
    sm_mhz = (r.get('clocks') or {}).get('sm_mhz') or 1965.0
    l2_bytes = roof_rd['l2_to_sm_gather_bytes'] + n_seq * rows * hw * 4
    l2_ceiling = 6300.0 * sm_mhz * 1e6 / 1e9
    roof_rd['l2_to_sm'] = dict(bytes=l2_bytes, achieved=l2_bytes / rd_t / 1e9, ceiling=l2_ceiling, unit='GB/s',
                               frac=l2_bytes / rd_t / 1e9 / l2_ceiling,
                               ceiling_source='B300_MICROARCH.md: LTS throughput cap ~6300 B/cycle (full chip) x sampled SM clock')
    roof_sel = dict(kernel='select_tc_kernel', bound='tensor', achieved=sel_flops / sel_t / 1e12, peak=pk['tflops'],
                    unit='TFLOP/s', frac=sel_flops / sel_t / 1e12 / pk['tflops'],
                    traffic=ncu_traffic(r['workload'], 'select_tc_kernel') if single else None,
                    us_per_launch=sel_t * 1e6, algorithmic_flops=sel_flops, executed_flop_multiplier=25.0 / 8.0,
                    executed_frac=25.0 / 8.0 * sel_flops / sel_t / 1e12 / pk['tflops'], peak_source=pk['source'] + ', burst')
    # whole step against SURVEY.md section 8d: FLOPs = 4 N HW CK + 2 nobj CV k HW; bytes as the reference would move them
    # (fp32, min(N, HW k) value rows) and as this design moves them (key image, bf16 shadow, unique rows)
    flops = n_seq * (4.0 * n_mem * hw * CK + 2.0 * rows * TOP_K * hw)
    bytes_8d = n_seq * 4.0 * (n_mem * (CK + 1) + 2 * CK * hw + rows * min(n_mem, hw * TOP_K) + rows * hw + 2 * n_mem)
    bytes_moved = n_seq * (n_mem * 544.0 + 2 * CK * hw * 4 + rows * uniq * val_bytes + rows * hw * 4 + 2 * n_mem * 4)
    t_tensor = flops / (pk['tflops_sustained'] * 1e12)
    step_s = r['ms_per_step'] / 1000.0
    roof_step = dict(flops=flops, bytes_section_8d=bytes_8d, bytes_moved=bytes_moved,
                     tensor_us=t_tensor * 1e6, hbm_us_section_8d=bytes_8d / (pk['hbm'] * 1e9) * 1e6,
                     hbm_us_moved=bytes_moved / (pk['hbm'] * 1e9) * 1e6,
                     roofline_us=max(t_tensor, bytes_8d / (pk['hbm'] * 1e9)) * 1e6, measured_us=step_s * 1e6,
                     frac=max(t_tensor, bytes_8d / (pk['hbm'] * 1e9)) / step_s,
                     frac_moved_bytes=max(t_tensor, bytes_moved / (pk['hbm'] * 1e9)) / step_s,
                     note='roofline_us = max(FLOPs / sustained bf16 peak, section-8d bytes / HBM peak); frac = roofline_us / measured_us')
    return roof_rd, roof_sel, roof_step


def stage_summary(r):
    return dict(pack_query=statistics.mean(r['pack_ms']) * 1e3, select=statistics.mean(r['sel_ms']) * 1e3,
                readout=statistics.mean(r['rd_ms']) * 1e3, step_median=statistics.median(r['step_ms']) * 1e3,
                step_kernel_by_kernel=r['step_kernel_by_kernel_us'])


def gpu_eager_baseline(ctx, workload, calls=10):
    """The reference's own torch op sequence (memory_util.py:24-27,46-54, memory_manager.py:55) on CUDA tensors on this
    GPU: the staged reference MemoryManager (oracle/_ref) -- or the oracle port -- event-timed.  SURVEY section 8d."""
    from tests import synth
    try:
        ref, kind, h, w, scale = reference_manager(workload, device=ctx.dev)
        g = torch.Generator().manual_seed(99)
        qk, qe = (t.to(ctx.dev) for t in synth.query(g, h, w))
        for _ in range(3):
            ref.match_memory(qk, qe)
        torch.cuda.synchronize()
        ms = []
        for i in range(calls):
            ctx.flush.fill_(i & 0xFF)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ref.match_memory(qk, qe)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        del ref
        torch.cuda.empty_cache()
        return dict(value=scale * 1000.0 / statistics.median(ms), unit=UNIT, ms_per_call=statistics.median(ms), kind=kind, calls=calls,
                    what='reference MemoryManager.match_memory with CUDA fp32 tensors (ATen / cuBLAS kernels, N x HW matrix '
                         'materialised), CUDA events, L2 flushed before every call')
    except Exception as exc:      # a baseline must never take the benchmark down
        return dict(unavailable=f'{type(exc).__name__}: {exc}'[:200])


def key_projection_times(ctx, h=30, w=54, reps=20):
    """SURVEY section 8f-4, the step right before the readout: vos_e_sam_b200.KeyProjection (one tcgen05 implicit GEMM for
    key / shrinkage / selection) against the reference module's three cuDNN convolutions (tracker/model/modules.py:194-211)
    on the same GPU; CUDA-graph replays, L2 flushed before each, median."""
    import statistics
    import torch.nn as nn
    import vos_e_sam_b200 as vos

    class RefKeyProjection(nn.Module):
        def __init__(self):
            super().__init__()
            self.key_proj, self.d_proj, self.e_proj = (nn.Conv2d(1024, 64, 3, padding=1), nn.Conv2d(1024, 1, 3, padding=1),
                                                       nn.Conv2d(1024, 64, 3, padding=1))

        def forward(self, x):
            return self.key_proj(x), self.d_proj(x) ** 2 + 1, torch.sigmoid(self.e_proj(x))

    def timed(fn):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn()
        ts = []
        for it in range(reps):
            ctx.flush.fill_(it & 0xff)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        return statistics.median(ts[3:])

    x = torch.randn(1, 1024, h, w, device=ctx.dev)
    ours, ref = vos.KeyProjection(1024, 64).to(ctx.dev), RefKeyProjection().to(ctx.dev)
    with torch.no_grad():
        t_ours = timed(lambda: ours(x, True, True))
        old = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        t_fp32 = timed(lambda: ref(x))
        torch.backends.cudnn.allow_tf32 = True
        t_tf32 = timed(lambda: ref(x))
        torch.backends.cudnn.allow_tf32 = old
    flops = 2.0 * h * w * 1024 * 9 * 129
    return dict(feature_map=[h, w], us=t_ours, algorithmic_tflops=flops / t_ours * 1e-6, executed_flop_multiplier=3,
                reference_module_cudnn_fp32_us=t_fp32, reference_module_cudnn_tf32_us=t_tf32,
                kernels='pack_x_kernel + keyproj_mma_kernel<144, 9> (tcgen05) + keyproj_finalize_kernel')


def maintenance_times(ctx, h=30, w=54, n_obj=5, steps=60):
    """What is NOT in the per-frame number: MemoryManager.add_memory (every mem_every-th frame) without / with the working
    -> long-term consolidation (memory_manager.py:152-190,211-286), and one eager match_memory call (Python + C call +
    kernels, device-synchronised wall clock) -- XMem's default bounds, DAVIS shape."""
    import statistics
    import vos_e_sam_b200 as vos
    from tests import synth
    cfg = xmem_config(max_mid_term_frames=10, min_mid_term_frames=5, max_long_term_elements=10000)
    m = vos.MemoryManager(cfg)
    g = torch.Generator().manual_seed(3)
    plain, consolidate, match = [], [], []
    for _ in range(steps):
        k, s, e = synth.keys(g, h * w)
        v = torch.randn(1, n_obj, CV, h, w, generator=g)
        a = (k.view(1, CK, h, w).to(ctx.dev), s.view(1, 1, h, w).to(ctx.dev), v.to(ctx.dev), list(range(1, n_obj + 1)))
        sel = e.view(1, CK, h, w).to(ctx.dev)
        before = m.work_mem.size
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.add_memory(*a, selection=sel)
        torch.cuda.synchronize()
        (consolidate if m.work_mem.size < before + h * w else plain).append((time.perf_counter() - t0) * 1e6)
        qk, qe = (t.to(ctx.dev) for t in synth.query(g, h, w))
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m.match_memory(qk, qe)
            torch.cuda.synchronize()
            match.append((time.perf_counter() - t0) * 1e6)
    med = lambda x: statistics.median(x) if x else None
    return dict(add_memory_us=med(plain), add_memory_with_consolidation_us=med(consolidate[1:] or consolidate),
                consolidations=len(consolidate), match_memory_eager_call_us=med(match),
                note='wall clock per call, device-synchronised on both sides (host launch work included); 5 objects, HW 1620')


def pcie_ceiling(ctx, nbytes, copies=20):
    """Raw concurrent D2H ceiling: every rank copies `nbytes` from HBM to pinned host memory `copies` times, no kernels.
    Returns aggregate GB/s (sum over ranks of bytes / slowest rank's time)."""
    src = torch.empty(nbytes, dtype=torch.uint8, device=ctx.dev)
    dst = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
    for i in range(3):
        dst[i % 2].copy_(src, non_blocking=True)
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(copies):
        dst[i % 2].copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    (dt_max,) = ctx.max_over_ranks(dt)
    ctx.barrier()
    return ctx.world * nbytes * copies / dt_max / 1e9


# ------------------------------------------------------------------------------------------------
def measure_sharded(ctx, K, W, exchange, shard, want_e2e, want_parity=True, gather=False, share=True):
    """BASELINE configs[3]: one LVOS-scale long-term bank sharded along N over the ranks (strong scaling).
    Every rank also holds the whole bank for the in-run 1-GPU unsharded reference time (rank 0 alone runs it)."""
    import vos_e_sam_b200 as vos
    from vos_e_sam_b200 import ops
    from vos_e_sam_b200.sharded import ShardedLongTermReadout
    from tests import synth
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    h, w, n_obj, n = LVOS['h'], LVOS['w'], LVOS['n_obj'], LVOS['n_long']
    hw, rows = h * w, n_obj * CV
    gl = torch.Generator().manual_seed(1234 + 4)     # same bank on every rank; each keeps its shard
    k, s, _ = synth.keys(gl, n)
    v = torch.randn(n_obj, CV, n, generator=gl)
    engine = ShardedLongTermReadout(xmem_config(vosmem_exchange=exchange, vosmem_shard=shard, vosmem_share_thresholds=share),
                                    rank, world, dev)
    engine.load_long_term(k, s, v)
    gq = torch.Generator().manual_seed(1234 + 40)
    pool = 4
    host_q = [tuple(t.pin_memory() for t in synth.query(gq, h, w)) for _ in range(pool)]
    dev_q = [(a.to(dev), b.to(dev)) for a, b in host_q]
    q_lo, q_hi = engine.query_range(hw)
    flush = ctx.flush

    def step(i, ev):
        qk, qe = dev_q[i % pool]
        flush.fill_(i & 0xFF)
        ev[0].record()
        engine.match(qk, qe, events=ev[1:4], gather=gather)    # after select + push / exchange + readout / gather
        ev[4].record()

    for i in range(W):
        step(i, new_events())
    ctx.barrier()
    events = [new_events() for _ in range(K)]
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(K):
        step(i, events[i])
    ctx.barrier()
    wall_s = time.perf_counter() - t0
    engine.check_status()
    step_ms = [e[0].elapsed_time(e[4]) for e in events]
    sel_ms = [e[0].elapsed_time(e[1]) for e in events]
    rd_ms = [e[1].elapsed_time(e[2]) for e in events]
    ga_ms = [e[2].elapsed_time(e[3]) for e in events]
    # device-timed per-step sum (max over ranks) and, because the ranks pipeline across frames (no barrier in the step),
    # the wall clock of the K back-to-back frames between two barriers (max over ranks)
    total_ms, wall_ms = ctx.max_over_ranks(sum(step_ms), wall_s * 1000.0)
    res = dict(value=K / (total_ms / 1000.0), ms_per_step=total_ms / K, frames_per_s_wall=K / (wall_ms / 1000.0),
               stage_us=dict(select_and_push=statistics.mean(sel_ms) * 1e3, exchange_and_readout=statistics.mean(rd_ms) * 1e3,
                             gather=statistics.mean(ga_ms) * 1e3, step_median=statistics.median(step_ms) * 1e3),
               exchange=exchange, shard=shard, gather_output=gather, scaling='strong',
               thresholds_shared_across_ranks=bool(share and exchange == 'peer' and world > 1),
               result='every rank holds the readout of its query slice' if not gather else 'full readout replicated on every rank',
               config=workload_config('lvos_sharded', world))

    # ---- end to end: pinned H2D of the query, D2H of this rank's slice (the ranks' slices together are the readout) ----
    if want_e2e:
        n_q = max(q_hi - q_lo, 1)
        host_outs = [torch.empty((rows, n_q), dtype=torch.float32).pin_memory() for _ in range(2)]
        copy_stream = torch.cuda.Stream()
        landed = [None, None]

        def e2e_step(i):
            a, b = host_q[i % pool]
            r = engine.match(a.to(dev, non_blocking=True), b.to(dev, non_blocking=True), gather=False)
            ready = torch.cuda.Event()
            ready.record()
            slot = i % 2
            if landed[slot] is not None:
                landed[slot].synchronize()
            copy_stream.wait_event(ready)
            with torch.cuda.stream(copy_stream):
                host_outs[slot][:, :r.shape[1]].copy_(r, non_blocking=True)
                r.record_stream(copy_stream)
                landed[slot] = torch.cuda.Event()
                landed[slot].record()
        for i in range(W):
            e2e_step(i)
        torch.cuda.synchronize()
        ctx.barrier()
        t0 = time.perf_counter()
        for i in range(K):
            e2e_step(i)
        for e in landed:
            if e is not None:
                e.synchronize()
        torch.cuda.synchronize()
        ctx.barrier()
        (e2e_ms,) = ctx.max_over_ranks((time.perf_counter() - t0) * 1000.0)
        res['e2e'] = dict(value=K / (e2e_ms / 1000.0), unit=UNIT, h2d_bytes_per_step=2 * CK * hw * 4,
                          d2h_bytes_per_step=rows * hw * 4, note='every rank copies its own query slice to the host')

    # ---- in-run 1-GPU unsharded reference (rank 0 alone) + parity of the sharded result against it ----
    t1_ms = 0.0
    parity = None
    if rank == 0:
        store = vos.KeyValueMemoryStore(count_usage=False)
        store.add(k.to(dev), [v.to(dev)], s.to(dev), None, None)
        segs, vals = [store.key_segment(0, n)], [store.value_segment(0, 0, with_usage=False)]
        out1 = torch.empty((rows, hw), dtype=torch.float32, device=dev)
        run1 = lambda j: ops.match(dev_q[j][0].flatten(2)[0], dev_q[j][1].flatten(2)[0], segs, vals, rows, TOP_K, out=out1)
        for i in range(W):
            run1(i % pool)
        torch.cuda.synchronize()
        ms = []
        for i in range(max(K // 2, 10)):
            flush.fill_(i & 0xFF)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run1(i % pool)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t1_ms = statistics.mean(ms)
    if world > 1:
        (t1_ms,) = ctx.max_over_ranks(t1_ms)
    if want_parity:
        # sharded (gathered on every rank through the engine's own gather path) vs the unsharded readout on rank 0
        full = engine.match(dev_q[0][0], dev_q[0][1], gather=True).clone()
        torch.cuda.synchronize()
        engine.check_status()
        if rank == 0:
            want = run1(0)
            torch.cuda.synchronize()
            diff = (full - want).abs().amax(0)                       # per query column
            scale = float(want.abs().max())
            # same kernels and operand bits on every path -> identical scores; a differing column means a different
            # survivor set (a swapped survivor moves ~1/30 of a query's weight, far above 1e-3)
            parity = dict(max_rel_err=float(diff.max()) / scale, columns_differing=int((diff > 1e-3 * scale).sum()), columns=hw,
                          sets_equal=bool((diff <= 1e-3 * scale).all()),
                          against='unsharded vosmem_match of the same bank on rank 0 (same frame)')
    res.update(n1_unsharded_ms=t1_ms, n1_unsharded_value=1000.0 / t1_ms if t1_ms else None,
               efficiency=(t1_ms / (world * res['ms_per_step'])) if t1_ms else None,
               efficiency_wall=(t1_ms / (world * wall_ms / K)) if t1_ms else None, parity=parity)
    return res


# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (there is no CPU path in vos_e_sam_b200)'
    torch.cuda.set_device(local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    ctx = Ctx(rank, world, local_rank)
    K, W = args.steps, args.warmup
    pk = peaks()

    if args.workload == 'lvos_sharded':
        # the sharded workload as the headline of this run (builder's sweeps); the default run carries it as `sharded`
        with ClockSampler(local_rank) as clocks:
            sh = measure_sharded(ctx, K, W, args.exchange, args.shard, want_e2e=True, gather=args.gather, share=not args.no_share)
        if rank == 0:
            line = dict(metric=METRIC, value=sh['value'], unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=sh['ms_per_step'],
                        higher_is_better=True, scaling='strong', vs_baseline=None, dtype='bf16 values, ~fp32 scores (bf16 hi/lo x3)',
                        data='synthetic', config=dict(sh['config'], exchange=args.exchange, shard=args.shard, gather=args.gather),
                        e2e=sh.get('e2e'), gpu_launches=K * (3 + (2 if args.gather and world > 1 else 0)), clocks=clocks.summary(),
                        stage_us=sh['stage_us'], sharded={k: v for k, v in sh.items() if k not in ('config', 'e2e')})
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    r = measure_dp(ctx, args.workload, K, W, want_e2e=True, sample_clocks=True)
    roof_rd, roof_sel, roof_step = rooflines(r, world, pk)
    ceiling = pcie_ceiling(ctx, r['n_seq'] * r['rows'] * r['hw'] * 4)
    extra = {}
    if args.workload == 'davis5' and not args.headline_only:
        Ks = max(20, min(K, 50))
        others = {}
        for wl in ('long_video', 'davis_batch'):
            o = measure_dp(ctx, wl, Ks, W, want_e2e=False, sample_clocks=False)
            o_rd, o_sel, o_step = rooflines(o, world, pk)
            others[wl] = dict(value=o['value'], unit=UNIT, ms_per_step=o['ms_per_step'], steps=Ks, stage_us=stage_summary(o),
                              config=workload_config(wl, world), roofline_step_frac=o_step['frac'],
                              select_frac_of_tensor_peak=o_sel['frac'], readout_frac_of_hbm_peak=o_rd['frac'])
            del o
            torch.cuda.empty_cache()
        extra['other_workloads'] = others
        sh = measure_sharded(ctx, Ks, W, args.exchange, 'n', want_e2e=True)
        extra['sharded'] = sh
        if rank == 0:
            extra['key_projection'] = key_projection_times(ctx)
            extra['maintenance'] = maintenance_times(ctx)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    dominant, other = (roof_rd, roof_sel) if roof_rd['us_per_launch'] >= roof_sel['us_per_launch'] else (roof_sel, roof_rd)
    cpu = eager = None
    if world == 1 and not args.no_cpu_baseline:
        rate, calls, threads, kind, _ = cpu_reference_rate(args.workload)
        cpu = dict(value=rate, unit=UNIT, cores=threads, kind=kind, sample=cpu_sample_text(kind, calls, args.workload))
        eager = gpu_eager_baseline(ctx, args.workload)
    d2h = r['n_seq'] * r['rows'] * r['hw'] * 4
    line = dict(metric=METRIC, value=r['value'], unit=UNIT, n_gpus=world, steps=K, warmup=W,
                ms_per_step=r['ms_per_step'], higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='bf16 values, ~fp32 scores (bf16 hi/lo x3)', data='synthetic',
                config=workload_config(args.workload, world),
                e2e=dict(value=r['e2e_value'], unit=UNIT, h2d_bytes_per_step=r['n_seq'] * 2 * CK * r['hw'] * 4, d2h_bytes_per_step=d2h,
                         pcie_ceiling=dict(d2h_gbs_all_ranks=ceiling, value=ceiling * 1e9 / d2h, unit=UNIT,
                                           how=f'{world} rank(s) x 20 pinned D2H copies of one frame\'s readout at once, no kernels'),
                         frac_of_pcie_ceiling=r['e2e_value'] / (ceiling * 1e9 / d2h)),
                gpu_launches=K * 2, clocks=r['clocks'], roofline=dominant, roofline_other=other, roofline_step=roof_step,
                cpu_baseline=cpu, gpu_eager_baseline=eager, stage_us=stage_summary(r), **extra)
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
    ap.add_argument('--headline-only', action='store_true', help='davis5: skip other_workloads and sharded')
    ap.add_argument('--shard', default='n', choices=['n', 'queries'],
                    help='lvos_sharded: shard the key axis (north_star) or, as a control, the query rows')
    ap.add_argument('--exchange', default='peer', choices=['nccl', 'peer'],
                    help='sharded bank: lists pushed into the owner\'s peer memory over NVLink, or one NCCL all-to-all')
    ap.add_argument('--gather', action='store_true', help='lvos_sharded: replicate the full readout on every rank')
    ap.add_argument('--no-share', action='store_true', help='lvos_sharded: no thresholds shared across the ranks (A/B)')
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
               '--workload', args.workload, '--exchange', args.exchange, '--shard', args.shard] + \
              (['--no-cpu-baseline'] if args.no_cpu_baseline else []) + (['--headline-only'] if args.headline_only else []) + \
              (['--gather'] if args.gather else []) + (['--no-share'] if args.no_share else [])
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
