#!/bin/bash
# A/B of N = 2048 kernel builds: tools/ab_2k.sh build/lib2k_a.so build/lib2k_b.so ...   (run under gpurun; 148 gates, n = 590, two parties)
# Every variant must pass the N = 2048 parity tests before its time counts.
for lib in "$@"; do
  export MKTFHE_B200_LIB=$PWD/$lib
  r=$(python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "n2048_parameter_sets" 2>&1 | tail -1)
  echo "$lib parity: $r"
  python tools/time_n2048.py 148 148 2>&1 | tail -1
done
