#!/bin/bash
# N-GPU visit of the FFT build (profiles/bench_r2_{q,t,u}_*): cross-engine tests, BASELINE configs[1] through one C-ABI context and through
# torchrun, configs[2] (4- and 8-party) through torchrun.   tools/gpu_multi_fft.sh <N> <tag>   (run under `gpurun --gpus N`)
N=$1; TAG=${2:-r2_t}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
python -m pytest tests/test_gpu_engines_agree.py -m gpu -x -q > $O/pytest_${TAG}_engines.log 2>&1; echo "engines rc=$?"; tail -2 $O/pytest_${TAG}_engines.log
python bench.py --gpus $N --abi-multi --steps 3 --latency-trials 5 --no-cpu-baseline > $O/bench_${TAG}_abimulti_${N}gpu.json 2> $O/bench_${TAG}_abimulti_${N}gpu.err; echo "abi-multi rc=$?"; cut -c1-120 $O/bench_${TAG}_abimulti_${N}gpu.json
$TR bench.py --gpus $N --steps 3 --latency-trials 5 > $O/bench_${TAG}_${N}gpu.json 2> $O/bench_${TAG}_${N}gpu.err; echo "torchrun rc=$?"; cut -c1-120 $O/bench_${TAG}_${N}gpu.json
for p in 4 8; do
  $TR bench.py --gpus $N --parties $p --steps 2 --warmup 3 --latency-trials 5 > $O/bench_${TAG}_${p}party_${N}gpu.json 2> $O/bench_${TAG}_${p}party_${N}gpu.err; echo "$p-party rc=$?"; cut -c1-120 $O/bench_${TAG}_${p}party_${N}gpu.json
done
