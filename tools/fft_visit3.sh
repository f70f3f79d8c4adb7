#!/bin/bash
# small batches: FFT channel with one six-warp gate per CTA against the NTT latency kernel (6 l warps per gate)
O=gpurun_out; mkdir -p $O
for p in 2 4 8; do for f in 0 1; do
  MKTFHE_B200_FFT_SMALL=$f python bench.py --parties $p --gates 148 --steps 3 --warmup 2 --no-cpu-baseline --latency-trials 20 > $O/fft_v3_${p}p_small$f.json 2> $O/fft_v3_${p}p_small$f.err
  python - "$p-party fft_small=$f" "$O/fft_v3_${p}p_small$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "148-gate launch ms", round(d["ms_per_step"], 3), "single bootstrap", d["ms_single_bootstrap_latency"], "dec", d.get("decryptions_correct"))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done; done
