#!/bin/bash
# 8-GPU evidence (run under gpurun --gpus 8): headline NAND metric, the encrypted conv layer sharded over the box, 4-party batches.
TAG=${1:-r1}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 > $O/bench_${TAG}_8gpu.json 2> $O/bench_${TAG}_8gpu.err; echo "nand rc=$?"; cut -c1-300 $O/bench_${TAG}_8gpu.json
$TR --master-port 29512 bench.py --gpus 8 --workload conv --steps 2 > $O/bench_${TAG}_conv_8gpu.json 2> $O/bench_${TAG}_conv_8gpu.err; echo "conv rc=$?"; cut -c1-400 $O/bench_${TAG}_conv_8gpu.json
$TR --master-port 29513 bench.py --gpus 8 --parties 4 --steps 2 --warmup 3 > $O/bench_${TAG}_4party_8gpu.json 2> $O/bench_${TAG}_4party_8gpu.err; echo "4party rc=$?"; cut -c1-300 $O/bench_${TAG}_4party_8gpu.json
