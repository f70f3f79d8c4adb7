#!/bin/bash
# last visit of the round: the whole GPU suite, smoke, and the default bench line (with the capture manifest of the shipped kernel in place)
TAG=${1:-r2_v}; O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_$TAG.log
python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$TAG.log
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; cut -c1-300 $O/bench_$TAG.json
