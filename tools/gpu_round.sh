#!/bin/bash
# One GPU-box visit for the round's evidence: tools/gpu_round.sh <tag>   (run under gpurun, one GPU)
#   1. pytest -m gpu, smoke()           -> gpurun_out/pytest_<tag>.log, smoke_<tag>.log
#   2. bench.py (both arms, no profiler) -> gpurun_out/bench_<tag>.json, bench_<tag>_reference.json
#   3. ncu launch list of a short bench  -> gpurun_out/launches_<tag>.csv      (gpu__time_duration only)
#   4. one ncu --set full capture of the dominant kernel -> gpurun_out/prof_<tag>.ncu-rep
# Numbers printed by the runs under ncu are never bench values.
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_$TAG.log
python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$TAG.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_${TAG}_reference.json 2> $O/bench_${TAG}_reference.err; echo "reference rc=$?"
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; cut -c1-400 $O/bench_$TAG.json
SHORT="python bench.py --gates 2368 --steps 2 --warmup 1 --no-cpu-baseline"
$SHORT > $O/short_$TAG.json 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv $SHORT > $O/ncu1_$TAG.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:blind_rotate -s 1 -c 1 -f -o $O/prof_$TAG $SHORT > $O/ncu2_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la $O | tail -12
