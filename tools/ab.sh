#!/bin/bash
# A/B harness for kernel experiments: tools/ab.sh <gates> <variant>...   (variants are build/lib_<variant>.so)
# Each variant first passes the golden NAND fixture (bit-exact), then is timed.
G=$1; shift
for v in "$@"; do
  MKTFHE_B200_LIB=$PWD/build/lib_$v.so python - "$v" "$G" <<'PY'
import json, subprocess, sys, os
v, G = sys.argv[1], sys.argv[2]
r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-q", "-m", "gpu", "-k", "golden or extprod or edge", "-x"], capture_output=True, text=True)
ok = "passed" in r.stdout.splitlines()[-1] and "failed" not in r.stdout.splitlines()[-1]
out = subprocess.run([sys.executable, "bench.py", "--gates", G, "--steps", "2", "--warmup", "2", "--no-cpu-baseline"], capture_output=True, text=True).stdout
d = json.loads(out.strip().splitlines()[-1])
print(f"{v:14s} parity={'OK' if ok else 'FAIL'} gates/s={d['value']:.0f} br_ms={d['roofline']['kernel_ms']:.1f} ks_ms={d['roofline']['keyswitch_ms']:.2f} dec={d['decryptions_correct']}", flush=True)
if not ok: print(r.stdout[-1500:])
PY
done
