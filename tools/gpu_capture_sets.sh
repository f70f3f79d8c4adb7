#!/bin/bash
# ncu --set full captures of the dominant kernel of the 4-party (l = 3) and 8-party (l = 4) sets on the FFT channel (one GPU)
#   -> gpurun_out/prof_<tag>_{4,8}party.ncu-rep ; summarised here with tools/ncu_summary.py + tools/ncu_manifest.py --engine fft64
TAG=${1:-r2_w}; O=gpurun_out; mkdir -p $O
for p in 4 8; do
  S="python bench.py --parties $p --gates 592 --steps 1 --warmup 1 --no-cpu-baseline --latency-trials 2"
  $S > $O/short_${TAG}_${p}party.json 2>&1 || continue
  ncu --set full --clock-control none --import-source on -k regex:blind_rotate_fft -s 1 -c 1 -f -o $O/prof_${TAG}_${p}party $S > $O/ncu_${TAG}_${p}party.log 2>&1; echo "$p-party ncu rc=$?"
done
ls -la $O | grep prof_${TAG}
