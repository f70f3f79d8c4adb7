#!/bin/bash
# 8-GPU visit of the FFT build, torchrun form only (run under `gpurun --gpus 8`)
N=$1; TAG=${2:-r2_u}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 3 --latency-trials 5 > $O/bench_${TAG}_${N}gpu.json 2> $O/bench_${TAG}_${N}gpu.err; echo "torchrun rc=$?"; cut -c1-120 $O/bench_${TAG}_${N}gpu.json
python bench.py --gpus $N --abi-multi --steps 3 --latency-trials 5 --no-cpu-baseline > $O/bench_${TAG}_abimulti_${N}gpu.json 2> $O/bench_${TAG}_abimulti_${N}gpu.err; echo "abi-multi rc=$?"; cut -c1-120 $O/bench_${TAG}_abimulti_${N}gpu.json
for p in 4 8; do
  $TR bench.py --gpus $N --parties $p --steps 2 --warmup 3 --latency-trials 5 > $O/bench_${TAG}_${p}party_${N}gpu.json 2> $O/bench_${TAG}_${p}party_${N}gpu.err; echo "$p-party rc=$?"; cut -c1-120 $O/bench_${TAG}_${p}party_${N}gpu.json
done
