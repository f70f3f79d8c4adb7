#!/usr/bin/env python
"""Manifest of one `ncu --set full` capture of the dominant kernel, read by bench.py instead of hard-coded constants:
   python tools/ncu_manifest.py profiles/ncu_r2_b_blind_rotate.txt --parties 2 --gates 2368 [--lib build/lib_x.so] -o profiles/ncu_capture_2party.json
The summary file comes from tools/ncu_summary.py; the manifest records its sha256 and the SASS identity (tools/kernel_id.py) of
the library the capture was taken from, so that a later kernel change invalidates it instead of silently going stale."""
import argparse
import hashlib
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kernel_id import hot_kernels, kernel_id  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("summary")
ap.add_argument("--parties", type=int, required=True)
ap.add_argument("--gates", type=int, required=True, help="gates in the captured launch")
ap.add_argument("--lib", default=None)
ap.add_argument("--N", type=int, default=1024)
ap.add_argument("--l", type=int, default=2)
ap.add_argument("--engine", default="ntt_rns", choices=["ntt_rns", "fft64"], help="which external-product kernels the capture is of")
ap.add_argument("-o", "--out", required=True)
a = ap.parse_args()
text = open(a.summary).read()
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "%": 1.0, "": 1.0}


def metric(name):
    m = re.search(rf"^{re.escape(name)} \[([^\]]*)\] = ([0-9.eE+-]+)", text, flags=re.M)
    if not m:
        raise SystemExit(f"{a.summary}: metric {name} not found")
    return float(m.group(2)) * UNIT.get(m.group(1), 1.0)


man = {
    "kernel": re.search(r"^== (.*?)  grid", text, flags=re.M).group(1),
    "engine": a.engine,
    "kernel_id": kernel_id(a.lib, hot_kernels(a.N, a.l, a.engine)),
    "kernel_id_covers": hot_kernels(a.N, a.l, a.engine),
    "parties": a.parties,
    "gates_in_launch": a.gates,
    "gpu_time_ms": metric("gpu__time_duration.sum"),
    "dram_bytes": metric("dram__bytes_read.sum") + metric("dram__bytes_write.sum"),
    "fmaheavy_pipe_busy": metric("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed") / 100.0,
    "fp64_pipe_busy": metric("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") / 100.0 if a.engine == "fft64" else None,
    "lsu_data_path_busy": metric("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") / 100.0 if a.engine == "fft64" else None,
    "l2_to_sm_bytes": metric("l1tex__m_xbar2l1tex_read_bytes.sum") if a.engine == "fft64" else None,
    "issue_active": metric("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
    "l2_hit_rate": metric("lts__t_sector_hit_rate.pct") / 100.0,
    "registers_per_thread": int(metric("launch__registers_per_thread")),
    "summary_file": os.path.relpath(a.summary, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
    "summary_sha256": hashlib.sha256(text.encode()).hexdigest(),
}
json.dump(man, open(a.out, "w"), indent=1)
print(json.dumps(man))
