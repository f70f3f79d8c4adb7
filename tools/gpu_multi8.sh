#!/bin/bash
# 8-GPU visit (run under `gpurun --gpus 8`), kept short -- it is charged eight-fold: tools/gpu_multi8.sh <tag>
#   the plain-C host and the Python binding over all 8 GPUs; 2-party through ONE C-ABI context; 8-party and the conv layer under torchrun
TAG=${1:-r2}; O=gpurun_out; mkdir -p $O; N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
python -m pytest tests/test_gpu_multi_and_circuits.py tests/test_cabi_direct.py -m gpu -x -q -k "multi or plain_c" > $O/pytest_${TAG}_multi_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_${TAG}_multi_${N}gpu.log
python bench.py --gpus $N --abi-multi --steps 3 --latency-trials 10 > $O/bench_${TAG}_abimulti_${N}gpu.json 2> $O/bench_${TAG}_abimulti_${N}gpu.err; echo "abi-multi rc=$?"; cut -c1-200 $O/bench_${TAG}_abimulti_${N}gpu.json
$TR bench.py --gpus $N --parties 8 --steps 2 --warmup 3 --latency-trials 5 > $O/bench_${TAG}_8party_${N}gpu.json 2> $O/bench_${TAG}_8party_${N}gpu.err; echo "8-party rc=$?"; cut -c1-200 $O/bench_${TAG}_8party_${N}gpu.json
$TR bench.py --gpus $N --workload conv --steps 2 > $O/bench_${TAG}_conv_${N}gpu.json 2> $O/bench_${TAG}_conv_${N}gpu.err; echo "conv rc=$?"; cut -c1-300 $O/bench_${TAG}_conv_${N}gpu.json
