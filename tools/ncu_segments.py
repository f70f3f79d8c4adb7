#!/usr/bin/env python
"""Split the SASS of a kernel at its barriers / back-branches and report, per segment, the share of warp time (stall samples),
executed warp instructions, fma-pipe slots and the top stall reasons:
   ncu -i X.ncu-rep --page source --csv > sass.csv ; python tools/ncu_segments.py sass.csv [gate_steps_in_launch]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
GS = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
FMA2 = ("IMAD.HI", "IMAD.WIDE")
segs, cur = [], None


def new_seg(label):
    global cur
    cur = dict(label=label, samples=0, inst=0, fma=0, alu=0, lsu=0, stalls=collections.Counter(), n=0, first=None)
    segs.append(cur)


new_seg("entry")
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[iS]
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    if not m:
        continue
    op = m.group(1)
    e, s = int(r[iE] or 0), int(r[iSamp] or 0)
    cur["samples"] += s; cur["inst"] += e; cur["n"] += 1
    if cur["first"] is None:
        cur["first"] = r[0]
    if op.startswith("IMAD") or op.startswith("UIMAD"):
        cur["fma"] += e * (2 if op.startswith(FMA2) else 1)
    elif op.startswith(("LDS", "STS", "LDG", "STG", "LD.", "ST.")):
        cur["lsu"] += e
    elif op.startswith(("IADD", "VIADD", "VIMNMX", "LOP3", "SHF", "LEA", "SEL", "PRMT", "ISETP", "MOV", "IABS", "FLO", "POPC")):
        cur["alu"] += e
    for i, name in stall_cols:
        cur["stalls"][name] += int(r[i] or 0)
    if op.startswith("BAR") or op.startswith("WARPSYNC") or (op == "BRA" and e > 0.5 * GS):
        new_seg(f"after {op} @{r[0][-5:]}")
tot = sum(s["samples"] for s in segs)
toti = sum(s["inst"] for s in segs)
print(f"samples {tot}  warp instructions {toti} = {toti / GS:.0f} per gate-step")
print(f"{'segment':34s} {'time%':>6s} {'inst/gs':>8s} {'fma/gs':>8s} {'alu/gs':>8s} {'lsu/gs':>7s} {'fma-util':>8s}  stalls")
for s in segs:
    if s["samples"] < 0.002 * tot and s["inst"] < 0.002 * toti:
        continue
    # fma-pipe utilisation estimate of the segment: its share of fma slots over its share of time, scaled by the kernel's measured total
    top = " ".join(f"{a}={100 * b / max(s['samples'], 1):.0f}%" for a, b in s["stalls"].most_common(4))
    print(f"{s['label'][:34]:34s} {100 * s['samples'] / tot:6.1f} {s['inst'] / GS:8.1f} {s['fma'] / GS:8.1f} {s['alu'] / GS:8.1f} {s['lsu'] / GS:7.1f} "
          f"{(s['fma'] / max(s['samples'], 1)) / (sum(x['fma'] for x in segs) / tot):8.2f}  {top}")
