#!/usr/bin/env python
"""Per-opcode executed-instruction mix and stall samples from an ncu source page:
   ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_opmix.py src.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops, samp = collections.Counter(), collections.Counter()
stall = collections.Counter()
tot = 0
for r in rows[2:]:
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
    if not m:
        continue
    op = m.group(1)
    e = int(r[iE] or 0)
    ops[op] += e
    samp[op] += int(r[iSamp] or 0)
    tot += e
    for i in stall_cols:
        stall[hdr[i]] += int(r[i] or 0)
print(f"total warp instructions executed: {tot}")
PIPE = {"IMAD": "fma", "IMAD.HI.U32": "fma x2", "IMAD.WIDE.U32": "fma x2", "IMAD.IADD": "fma", "IMAD.MOV.U32": "fma", "IMAD.SHL.U32": "fma", "IMAD.X": "fma", "IMAD.MOV": "fma",
        "IADD3": "alu", "VIADDMNMX.U32": "alu", "VIMNMX.U32": "alu", "LOP3.LUT": "alu", "SHF.R.U32.HI": "alu", "SHF.L.U32": "alu", "LEA": "alu", "SEL": "alu", "MOV": "alu",
        "ISETP.NE.AND": "alu", "VIADD": "alu", "IADD3.X": "alu", "PRMT": "alu", "SHF.R.U64": "alu", "SHF.L.U64.HI": "alu", "LEA.HI.X": "alu", "LEA.HI": "alu"}
pipes = collections.Counter()
for op, e in ops.most_common(40):
    print(f"{op:22s} {e:14d} {100.0 * e / tot:6.2f}%  samples {samp[op]:8d}  {PIPE.get(op, '')}")
for op, e in ops.items():
    pipes[PIPE.get(op, "other").split()[0]] += e * (2 if "x2" in PIPE.get(op, "") else 1)
print("pipe slots (x2 for half-rate):", dict(pipes))
ts = sum(stall.values())
print("stall samples:", {k: f"{100.0 * v / ts:.1f}%" for k, v in stall.most_common(10)})
