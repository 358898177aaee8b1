#!/usr/bin/env python
"""Print selected raw metrics of every kernel in an .ncu-rep (read here, on the CPU box).

    python scripts/ncu_metrics.py gpurun_out/prof.ncu-rep [kernel-substring] [--all]

Default selection: duration, DRAM / L2 / L1 traffic, tensor-pipe and issue utilisation, warp-stall breakdown.
"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct',
        'lts__t_bytes.sum', 'lts__throughput.avg.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_l1tex2xbar_write_bytes.sum', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__throughput.avg.pct', 'sm__pipe_tensor_cycles_active.avg.pct', 'sm__warps_active.avg.pct',
        'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit',
        'smsp__average_warps_issue_stalled', 'smsp__average_warp_latency_per_inst_issued']
SKIP = ['per_second', 'evict', 'atom', 'op_red', 'peak_sustained ', '.max.', '.min.', '.sum.pct']


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith('--') else ''
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index('Kernel Name')
    idx = [i for i, h in enumerate(hdr) if any(w in h for w in WANT) and not any(s in h for s in SKIP)]
    for r in rows[2:]:
        if sub and sub not in r[name_col]:
            continue
        print(r[name_col][:100])
        for i in idx:
            if r[i] not in ('', 'n/a'):
                print(f'    {hdr[i]} [{units[i]}] = {r[i]}')


if __name__ == '__main__':
    main()
