#!/bin/bash
# first GPU visit of the FP64 FFT channel: parity subset, then A/B timing against the NTT kernels
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "golden or extprod or edge or gates_bit_exact" > $O/fft_v1_parity.log 2>&1; echo "parity rc=$?"; tail -5 $O/fft_v1_parity.log
for f in 1 0; do
  MKTFHE_B200_FFT=$f python bench.py --gates 16384 --steps 2 --warmup 1 --no-cpu-baseline --latency-trials 2 > $O/fft_v1_bench_fft$f.json 2> $O/fft_v1_bench_fft$f.err; echo "bench fft=$f rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$O/fft_v1_bench_fft$f.json").read().strip().splitlines()[-1])
    print("fft=$f gates/s", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "dec", d.get("decryptions_correct"), "clocks", d.get("clocks"))
except Exception as e:
    print("parse failed", e); print(open("$O/fft_v1_bench_fft$f.err").read()[-2000:])
PY
done
