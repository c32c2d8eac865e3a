#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show what the built library actually uses on sm_100a (tcgen05 MMA / TMEM
loads / bulk-TMA copies / mbarriers / warp-match / REDUX ...), from `cuobjdump -sass` of the in-tree libanqs_b200.so.

    python scripts/sass_summary.py > profiles/sass_summary.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'anqs_quantum_chemistry_b200', 'libanqs_b200.so')
MNEMONICS = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTCCP', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'MATCH', 'REDUX', 'POPC',
             'DMMA', 'DFMA', 'DADD', 'LDS', 'STS', 'LDG', 'STG', 'ATOMG', 'ATOMS', 'SHFL', 'VOTE', 'BAR']
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r'arch = (sm_\w+)', out)))
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r'\(.*', '', kern).replace('void ', '').replace('anqs::', '')
        counts.setdefault(kern, collections.Counter())
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and kern:
        op = m.group(1)
        counts[kern]['_all'] += 1
        for mn in MNEMONICS:
            if op.startswith(mn):
                counts[kern][mn] += 1
                total[mn] += 1
print(f'# SASS summary of {os.path.relpath(so, ROOT)} (cuobjdump -sass; architectures in the fatbin: {", ".join(arch)})')
print('# mma.sync.m8n8k4.f64 -> DMMA (FP64 tensor-core path of the float64 network kernels);')
print('# tcgen05.mma -> UTCHMMA (tf32/f16 kinds), tcgen05.ld -> LDTM, tcgen05.commit / mbarrier -> UTCBAR / SYNCS, cp.async.bulk -> UBLKCP,')
print('# __match_any_sync -> MATCH, __reduce_*_sync -> REDUX.  Counts are static instruction counts per kernel (all template instances summed).')
used = [mn for mn in MNEMONICS if total[mn]]
print(f'{"kernel":48s} {"instr":>7s} ' + ' '.join(f'{mn:>7s}' for mn in used))
for k, c in counts.items():
    print(f'{k[:48]:48s} {c["_all"]:7d} ' + ' '.join(f'{c[mn]:7d}' for mn in used))
print(f'{"TOTAL":48s} {sum(c["_all"] for c in counts.values()):7d} ' + ' '.join(f'{total[mn]:7d}' for mn in used))
