#!/bin/bash
# ncu --set full captures of the dominant kernel of the other parameter sets (one GPU): 4-party (l = 3), 8-party (l = 4), 16-party (N = 2048)
#   -> gpurun_out/prof_<tag>_{4,8,16}party.ncu-rep ; summarised here with tools/ncu_summary.py + tools/ncu_manifest.py
TAG=${1:-r2}; O=gpurun_out; mkdir -p $O
for p in 4 8; do
  S="python bench.py --parties $p --gates 592 --steps 1 --warmup 1 --no-cpu-baseline --latency-trials 2"
  $S > $O/short_${TAG}_${p}party.json 2>&1 || continue
  ncu --set full --clock-control none --import-source on -k regex:blind_rotate_kernel -s 1 -c 1 -f -o $O/prof_${TAG}_${p}party $S > $O/ncu_${TAG}_${p}party.log 2>&1; echo "$p-party ncu rc=$?"
done
S="python bench.py --parties 16 --gates 148 --steps 1 --warmup 1 --no-cpu-baseline --latency-trials 2"
$S > $O/short_${TAG}_16party.json 2>&1 && ncu --set full --clock-control none --import-source on -k regex:blind_rotate2k16 -s 1 -c 1 -f -o $O/prof_${TAG}_16party $S > $O/ncu_${TAG}_16party.log 2>&1; echo "16-party ncu rc=$?"
ls -la $O | grep prof_${TAG}
