#!/bin/bash
# ncu --set full capture of the FFT-channel blind rotation (one GPU): tools/fft_capture.sh <tag>
TAG=${1:-fft}; O=gpurun_out; mkdir -p $O
SHORT="python bench.py --gates 2368 --steps 2 --warmup 1 --no-cpu-baseline --latency-trials 2"
$SHORT > $O/short_$TAG.json 2>&1 || { tail -20 $O/short_$TAG.json; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_fft -s 1 -c 1 -f -o $O/prof_$TAG $SHORT > $O/ncu_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la $O | grep prof_$TAG
