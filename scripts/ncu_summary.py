"""Summarise an .ncu-rep: per-kernel headline metrics + SASS hot spots.  usage: ncu_summary.py rep [kernel-regex]"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]; rx = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]; idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for r in rows[2:]:
    print(r[idx['Kernel Name']][:100])
    for w in want:
        if w in idx: print(f'    {w} [{units[idx[w]]}] = {r[idx[w]]}')
if rx:
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
    recs = []; ts = ti = 0
    for r in rows[2:]:
        if len(r) != len(hdr): break
        try: s_ = int(r[idx['# Samples']]); ie = int(r[idx['Instructions Executed']])
        except ValueError: continue
        recs.append((r[idx['Address']], s_, ie, r[idx['Source']])); ts += s_; ti += ie
    print(f'--- {rx}: samples {ts}, warp instructions {ti}')
    byop = collections.defaultdict(lambda: [0, 0])
    for a, s_, ie, srcl in recs:
        t = srcl.split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        byop[op][0] += s_; byop[op][1] += ie
    for op, (s_, ie) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:16]:
        print(f'    {op:10s} samples {100 * s_ / max(ts, 1):5.1f}%   instr {ie:9d} {100 * ie / max(ti, 1):5.1f}%')
    print('    hottest SASS lines:')
    for a, s_, ie, srcl in sorted(recs, key=lambda x: -x[1])[:18]:
        print(f'    {a[-5:]} {s_:5d} {ie:8d}  {srcl[:80]}')
