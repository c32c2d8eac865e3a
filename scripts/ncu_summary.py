#!/usr/bin/env python
"""Summarises an .ncu-rep (read here, no GPU needed): key raw metrics + the hottest source lines.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25] > profiles/<name>.txt
"""
import csv
import io
import re
import subprocess
import sys

KEYS = (r'^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct|lts__t_sectors\.sum|'
        r'lts__t_sector_hit_rate|lts__throughput\.avg\.pct|l1tex__t_sector_hit_rate|'
        r'l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__throughput\.avg\.pct|sm__issue_active\.avg\.pct|'
        r'sm__inst_executed_pipe_(alu|fma|xu|lsu|fp64|uniform)\.avg\.pct_of_peak_sustained_active|'
        r'sm__pipe_tensor.*cycles_active\.avg\.pct|smsp__inst_executed\.sum$|sm__warps_active\.avg\.pct|'
        r'launch__registers_per_thread$|launch__shared_mem_per_block_(dynamic|static)|launch__grid_size|launch__block_size|'
        r'launch__occupancy_limit|smsp__average_warps_issue_stalled_.*_per_issue_active|'
        r'l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|l1tex__t_requests_pipe_lsu_mem_global_op_ld\.sum$|'
        r'l1tex__t_sectors_pipe_lsu_mem_global_op_ld\.sum$|smsp__sass_inst_executed_op_shared)')


def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 25
    rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
        print(f'== kernel: {name}')
        for h, u, v in zip(hdr, units, r):
            if re.search(KEYS, h) and v not in ('', '0'):
                print(f'  {h:95s} {v} {u}')
    src = run(['-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'])
    lines, fname, hdr = [], '?', None
    for r in csv.reader(io.StringIO(src)):
        if not r:
            continue
        if r[0] == 'File Path':
            fname = r[1].split('/')[-1]
        elif r[0] == 'Line No':
            hdr = r
        elif hdr and r[0].isdigit():
            try:
                samples = float(r[hdr.index('# Samples')])
                inst = float(r[hdr.index('Instructions Executed')])
            except Exception:
                continue
            lines.append((samples, inst, f'{fname}:{r[0]}', r[1].strip()))
    tot_s = sum(x[0] for x in lines) or 1.0
    tot_i = sum(x[1] for x in lines) or 1.0
    print(f'== hottest source lines (stall samples total {tot_s:.0f}, warp instructions total {tot_i:.4g})')
    print('   samples%  inst%   where   source')
    for smp, inst, where, text in sorted(lines, key=lambda x: -x[0])[:top]:
        print(f'  {100 * smp / tot_s:6.1f}% {100 * inst / tot_i:6.1f}%  {where:22s} {text[:120]}')


if __name__ == '__main__':
    main()
