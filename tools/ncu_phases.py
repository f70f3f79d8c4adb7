#!/usr/bin/env python
"""Phase table of blind_rotate_kernel<2,2,6> from the SASS page of an ncu capture.  Phases are delimited by the kernel's own
synchronisation instructions in the main loop (gate / pair barriers, warp syncs), found by their execution counts:
   ncu -i X.ncu-rep --page source --csv > sass.csv ; python tools/ncu_phases.py sass.csv <gate_steps_in_launch> <fmaheavy_busy_pct>"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1]))]
GS = float(sys.argv[2])
BUSY = float(sys.argv[3]) if len(sys.argv) > 3 else None
hdr = rows[1]
body = [r for r in rows[2:] if len(r) >= len(hdr)]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
# binding math pipe: IMAD for the NTT kernels; "fp64" as 4th argument counts DFMA / DADD / DMUL instead (FFT-channel kernels)
MATH = ("DFMA", "DADD", "DMUL") if len(sys.argv) > 4 and sys.argv[4] == "fp64" else ("IMAD", "UIMAD")
stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]


def opcode(src):
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    return m.group(1) if m else ""


# main-loop boundaries: sync instructions executed about once (or l times) per warp per step
marks = []
for i, r in enumerate(body):
    op = opcode(r[iS])
    e = int(r[iE] or 0) / GS
    if e > 3 and (op.startswith("BAR") or op.startswith("WARPSYNC") or op == "BRA.DIV" or (op == "BRA" and "@" in r[iS])):
        marks.append((i, op, e))
segs = []
prev = marks[0][0] if marks else 0
for (i, op, e) in marks[1:]:
    segs.append((prev, i, op, e))
    prev = i
tot_s = sum(int(r[iSamp] or 0) for r in body)
tot_f = 0
out = []
for (a, b, op, e) in segs:
    d = dict(samples=0, inst=0, fma=0, alu=0, lsu=0, stalls=collections.Counter())
    for r in body[a + 1:b + 1]:
        o = opcode(r[iS]); ex = int(r[iE] or 0); d["samples"] += int(r[iSamp] or 0); d["inst"] += ex
        if o.startswith(MATH):
            d["fma"] += ex * (2 if o.startswith(("IMAD.HI", "IMAD.WIDE")) else 1)
        elif o.startswith(("LDS", "STS", "LDG", "STG")):
            d["lsu"] += ex
        elif o.startswith(("IADD", "VIADD", "VIMNMX", "LOP3", "SHF", "LEA", "SEL", "PRMT", "ISETP", "MOV", "CS2R")):
            d["alu"] += ex
        for i, name in stall_cols:
            d["stalls"][name] += int(r[i] or 0)
    tot_f += d["fma"]
    out.append((a, b, op, e, d))
allf = sum(int(r[iE] or 0) * (2 if opcode(r[iS]).startswith(("IMAD.HI", "IMAD.WIDE")) else 1) for r in body if opcode(r[iS]).startswith(MATH))
print(f"samples {tot_s}; fma slots {allf / GS:.0f} per gate-step (main loop {tot_f / GS:.0f}); fmaheavy busy {BUSY} %")
print(f"{'sass idx':>11s} {'ends with':22s} {'time%':>6s} {'inst/gs':>8s} {'fma/gs':>8s} {'alu/gs':>8s} {'lsu/gs':>7s} {'pipe%':>6s}  stalls")
for (a, b, op, e, d) in out:
    util = (d["fma"] / max(d["samples"], 1)) / (allf / tot_s) * (BUSY or 100.0)
    top = " ".join(f"{x}={100 * y / max(d['samples'], 1):.0f}%" for x, y in d["stalls"].most_common(5))
    print(f"{a:5d}-{b:5d} {op[:22]:22s} {100 * d['samples'] / tot_s:6.1f} {d['inst'] / GS:8.1f} {d['fma'] / GS:8.1f} {d['alu'] / GS:8.1f} {d['lsu'] / GS:7.1f} {util:6.1f}  {top}")
