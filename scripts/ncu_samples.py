#!/usr/bin/env python
"""Stall-sample share per source line (and the dominant stall reasons) of one kernel launch of an .ncu-rep.
    python scripts/ncu_samples.py rep.ncu-rep <kernel regex> [--top N] [--launch I]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 30
skip = int(sys.argv[sys.argv.index('--launch') + 1]) if '--launch' in sys.argv else 0
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '-k', 'regex:' + pat,
                      '--launch-skip', str(skip), '--launch-count', '1'], capture_output=True, text=True).stdout
lines, fname, hdr = {}, '?', None
reasons = {}
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = r
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
    elif hdr and r[0].isdigit():
        try:
            smp = float(r[hdr.index('# Samples')]); inst = float(r[hdr.index('Instructions Executed')])
        except Exception:
            continue
        k = (fname, int(r[0]))
        a = lines.setdefault(k, [0.0, 0.0, r[1].strip(), {}])
        a[0] += smp; a[1] += inst
        for i in stall_cols:
            try:
                v = float(r[i])
            except Exception:
                continue
            a[3][hdr[i]] = a[3].get(hdr[i], 0.0) + v
            reasons[hdr[i]] = reasons.get(hdr[i], 0.0) + v
ts = sum(v[0] for v in lines.values()) or 1; ti = sum(v[1] for v in lines.values()) or 1
tr = sum(reasons.values()) or 1
print(f'total warp instructions {ti:.4g}, stall samples {ts:.0f}')
print('stall reasons: ' + ', '.join(f'{k[6:]} {100*v/tr:.1f}%' for k, v in sorted(reasons.items(), key=lambda kv: -kv[1])[:8]))
for (f, ln), (smp, inst, text, rs) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    best = ', '.join(f'{k[6:]} {100*v/max(sum(rs.values()),1):.0f}%' for k, v in sorted(rs.items(), key=lambda kv: -kv[1])[:2])
    print(f'{100*smp/ts:6.2f}% smp {100*inst/ti:6.2f}% inst  {f}:{ln:<4d} [{best}] {text[:80]}')
