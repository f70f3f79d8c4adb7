#!/bin/bash
# FFT channel as the default throughput kernel: full GPU test suite, then FFT-vs-NTT timing on the other parameter sets
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_fft_v2.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_fft_v2.log
run() { # tag env... -- bench args
  tag=$1; shift; e=""; while [ "$1" != "--" ]; do e="$e $1"; shift; done; shift
  env $e python bench.py "$@" --steps 2 --warmup 1 --no-cpu-baseline --latency-trials 2 > $O/fft_v2_$tag.json 2> $O/fft_v2_$tag.err
  python - "$tag" "$O/fft_v2_$tag.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "gates/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "dec", d.get("decryptions_correct"))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
run 4p_fft MKTFHE_B200_FFT=1 -- --parties 4 --gates 4736
run 4p_ntt MKTFHE_B200_FFT=0 -- --parties 4 --gates 4736
run 8p_fft MKTFHE_B200_FFT=1 -- --parties 8 --gates 2368
run 8p_ntt MKTFHE_B200_FFT=0 -- --parties 8 --gates 2368
run single_fft MKTFHE_B200_FFT=1 -- --workload single --gates 16384
run single_ntt MKTFHE_B200_FFT=0 -- --workload single --gates 16384
