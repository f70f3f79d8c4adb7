#!/bin/bash
# tools/fft_ab_capture.sh <variant>...: tools/ab.sh on every variant, then one ncu --set full capture of each variant's FFT blind rotation
O=gpurun_out; mkdir -p $O
tools/ab.sh 16384 "$@"
SHORT="python bench.py --gates 2368 --steps 2 --warmup 1 --no-cpu-baseline --latency-trials 2"
for v in "$@"; do
  MKTFHE_B200_LIB=$PWD/build/lib_$v.so ncu --set full --clock-control none --import-source on -k regex:blind_rotate_fft -s 1 -c 1 -f -o $O/prof_$v $SHORT > $O/ncu_$v.log 2>&1; echo "ncu $v rc=$?"
done
