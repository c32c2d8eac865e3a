#!/usr/bin/env python
"""Instruction and stall-sample share per source line of one kernel of an .ncu-rep, sorted by instructions.
    python scripts/ncu_lines.py rep.ncu-rep <kernel regex> [--top N]"""
import csv, io, re, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 40
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '-k', 'regex:' + pat, '--launch-count', '1'],
                     capture_output=True, text=True).stdout
lines, fname, hdr = {}, '?', None
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = r
    elif hdr and r[0].isdigit():
        try:
            smp = float(r[hdr.index('# Samples')]); inst = float(r[hdr.index('Instructions Executed')])
        except Exception:
            continue
        k = (fname, int(r[0]))
        a = lines.setdefault(k, [0.0, 0.0, r[1].strip()])
        a[0] += smp; a[1] += inst
ti = sum(v[1] for v in lines.values()) or 1; ts = sum(v[0] for v in lines.values()) or 1
print(f'total warp instructions {ti:.4g}, stall samples {ts:.0f}')
for (f, ln), (smp, inst, text) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{100*inst/ti:6.2f}% inst {100*smp/ts:6.2f}% smp  {f}:{ln:<5d} {text[:110]}')
