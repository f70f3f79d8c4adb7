#!/bin/bash
# A/B of run-time knobs of the shipped library: tools/ab_env.sh <gates> "VAR=val ..." ["VAR=val ..."]...   ("-" = no override)
G=$1; shift
for e in "$@"; do
  [ "$e" = "-" ] && e=""
  env $e python - "$e" "$G" <<'PY'
import json, subprocess, sys
e, G = sys.argv[1], sys.argv[2]
r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-q", "-m", "gpu", "-k", "golden or extprod or edge or full_batch or gates_bit_exact", "-x"], capture_output=True, text=True)
last = r.stdout.strip().splitlines()[-1]
ok = "passed" in last and "failed" not in last
out = subprocess.run([sys.executable, "bench.py", "--gates", G, "--steps", "2", "--warmup", "2", "--no-cpu-baseline"], capture_output=True, text=True).stdout
d = json.loads(out.strip().splitlines()[-1])
print(f"{e or 'default':28s} parity={'OK' if ok else 'FAIL'} gates={G} gates/s={d['value']:.0f} br_ms={d['roofline']['kernel_ms']:.2f} dec={d['decryptions_correct']}", flush=True)
if not ok: print(r.stdout[-1500:])
PY
done
