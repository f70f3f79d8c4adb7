#!/usr/bin/env python
"""Per-source-line samples / instructions / top stalls from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > f.csv`:
   python tools/ncu_lines.py f.csv [gate_steps_in_launch] [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
GS = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
TOP = int(sys.argv[3]) if len(sys.argv) > 3 else 70
hdr, fname = None, None
per_line = collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        fname = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name':
        continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
        continue
    if len(r) > len(hdr):          # a source line with unescaped quotes/commas split into extra cells: realign from the right
        extra = len(r) - len(hdr)
        r = [r[0], ",".join(r[1:2 + extra])] + r[2 + extra:]
    iS, iE = hdr.index('# Samples'), hdr.index('Instructions Executed')
    key = (fname, int(r[0]))
    d = per_line.setdefault(key, dict(src=r[1].strip(), samples=0, inst=0, stalls=collections.Counter()))
    d['samples'] += int(r[iS] or 0); d['inst'] += int(r[iE] or 0)
    for i, h in enumerate(hdr):
        if h.startswith('stall_') and 'Not Issued' not in h:
            d['stalls'][h[6:]] += int(r[i] or 0)
tot = sum(v['samples'] for v in per_line.values())
toti = sum(v['inst'] for v in per_line.values())
print(f"total samples {tot}, warp instructions {toti} ({toti / GS:.0f} per gate-step), lines {len(per_line)}")
byfile = collections.Counter()
for k, v in per_line.items():
    byfile[k[0]] += v['samples']
print({k: f"{100 * v / tot:.1f}%" for k, v in byfile.items()})
for k, v in sorted(per_line.items(), key=lambda kv: -kv[1]['samples'])[:TOP]:
    top = v['stalls'].most_common(4)
    print(f"{k[0]}:{k[1]:4d} {100 * v['samples'] / tot:5.1f}% inst/gs {v['inst'] / GS:7.1f} spi {v['samples'] / max(v['inst'], 1) * toti / tot:5.2f}  "
          f"{' '.join(f'{a}={100 * b / max(v['samples'], 1):.0f}%' for a, b in top)} | {v['src'][:90]}")
