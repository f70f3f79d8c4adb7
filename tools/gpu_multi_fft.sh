#!/bin/bash
# 2-GPU check of the FFT build: multi-device tests, one C-ABI context over N GPUs, torchrun form   (run under `gpurun --gpus N`)
N=$1; TAG=${2:-r2_q}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
python -m pytest tests/test_gpu_multi_and_circuits.py tests/test_cabi_direct.py -m gpu -x -q > $O/pytest_${TAG}_multi_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_${TAG}_multi_${N}gpu.log
python bench.py --gpus $N --abi-multi --steps 3 --latency-trials 5 --no-cpu-baseline > $O/bench_${TAG}_abimulti_${N}gpu.json 2> $O/bench_${TAG}_abimulti_${N}gpu.err; echo "abi-multi rc=$?"; cut -c1-200 $O/bench_${TAG}_abimulti_${N}gpu.json
$TR bench.py --gpus $N --steps 3 --latency-trials 5 > $O/bench_${TAG}_${N}gpu.json 2> $O/bench_${TAG}_${N}gpu.err; echo "torchrun rc=$?"; cut -c1-200 $O/bench_${TAG}_${N}gpu.json
