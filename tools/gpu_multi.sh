#!/bin/bash
# Multi-GPU visit: tools/gpu_multi.sh <N> <tag>   (run under `gpurun --gpus N`)
#   1. the multi-device C-ABI tests (pytest + the plain-C host over all N GPUs)
#   2. 2-party: one process / one context over N GPUs (--abi-multi) beside the torchrun form  -> bench_<tag>_abimulti_<N>gpu.json, bench_<tag>_<N>gpu.json
#   3. BASELINE configs[2]: 4- and 8-party batches on N GPUs (torchrun)                       -> bench_<tag>_{4,8}party_<N>gpu.json
N=$1; TAG=${2:-r2}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
python -m pytest tests/test_gpu_multi_and_circuits.py tests/test_cabi_direct.py -m gpu -x -q > $O/pytest_${TAG}_multi_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_${TAG}_multi_${N}gpu.log
python bench.py --gpus $N --abi-multi > $O/bench_${TAG}_abimulti_${N}gpu.json 2> $O/bench_${TAG}_abimulti_${N}gpu.err; echo "abi-multi rc=$?"; cut -c1-260 $O/bench_${TAG}_abimulti_${N}gpu.json
[ "$N" -le 4 ] && MKTFHE_B200_BCAST=nccl python bench.py --gpus $N --abi-multi --steps 2 --latency-trials 5 > $O/bench_${TAG}_abimulti_nccl_${N}gpu.json 2> $O/bench_${TAG}_abimulti_nccl_${N}gpu.err; echo "abi-multi nccl rc=$?"; cut -c1-200 $O/bench_${TAG}_abimulti_nccl_${N}gpu.json
$TR bench.py --gpus $N > $O/bench_${TAG}_${N}gpu.json 2> $O/bench_${TAG}_${N}gpu.err; echo "torchrun rc=$?"; cut -c1-260 $O/bench_${TAG}_${N}gpu.json
for p in ${PARTIES:-4 8}; do
  $TR bench.py --gpus $N --parties $p --steps 3 --warmup 3 --latency-trials 20 > $O/bench_${TAG}_${p}party_${N}gpu.json 2> $O/bench_${TAG}_${p}party_${N}gpu.err; echo "$p-party rc=$?"; cut -c1-200 $O/bench_${TAG}_${p}party_${N}gpu.json
done
