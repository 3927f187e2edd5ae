#!/usr/bin/env python
"""Summarise ncu outputs into text for profiles/.

  python tools/ncu_summary.py launches gpurun_out/x_launches.csv      # per-kernel totals and shares
  python tools/ncu_summary.py rep gpurun_out/x.ncu-rep                # key metrics of every captured launch
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
    'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__cycles_elapsed.max', 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = defaultdict(lambda: [0, 0.0])
    scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}
    for r in rows[1:]:
        v = float(r[vi].replace(',', '')) * scale.get(r[ui], 1.0)
        agg[r[ki]][0] += 1
        agg[r[ki]][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f'# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.3f} ms total (cold-cache, serialised)')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{v[1]:10.3f} ms {100 * v[1] / tot:5.1f}%  x{v[0]:<4d} avg {v[1] / v[0] * 1e3:9.1f} us  {k[:110]}')


def rep(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index('Kernel Name')
    for r in rows[2:]:
        print(f'## {r[kn][:120]}')
        for mname in METRICS:
            if mname in hdr:
                i = hdr.index(mname)
                print(f'  {mname:82s} {r[i]:>18s} {units[i]}')
        i1, i2 = hdr.index('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum'), hdr.index(
            'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum')
        try:
            print(f"  {'sectors/request (global loads)':82s} {float(r[i1]) / float(r[i2]):18.2f}")
        except (ValueError, ZeroDivisionError):
            pass


if __name__ == '__main__':
    {'launches': launches, 'rep': rep}[sys.argv[1]](sys.argv[2])
