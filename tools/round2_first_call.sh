#!/bin/bash
# First GPU-box visit of round 2 (run under gpurun, one GPU):   tools/round2_first_call.sh
#   1. pytest -m gpu (includes tests/test_zz_gpu_path_internals.py, written in round 1 after the GPU budget was spent)
#   2. the FP64-pipe and IMMA probes that decide DESIGN.md section 7 item 3    -> gpurun_out/fp64_pipe_ubench_r2.txt, imma_probe_ubench_r2.txt
#   3. circuits with host-resident vs HBM-resident ciphertexts (adder, less)    -> gpurun_out/circuits_r2.txt
#   4. the default bench line                                                    -> gpurun_out/bench_r2_a.json
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_r2_a.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_r2_a.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_pipe_ubench tools/fp64_pipe_ubench.cu && tools/fp64_pipe_ubench | tee $O/fp64_pipe_ubench_r2.txt
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/imma_probe_ubench tools/imma_probe_ubench.cu && tools/imma_probe_ubench | tee $O/imma_probe_ubench_r2.txt
for w in adder less; do
  for r in "" "--resident"; do
    python bench.py --workload $w $r --steps 2 2>&1 | tail -1 | cut -c1-400 | tee -a $O/circuits_r2.txt
  done
done
python bench.py > $O/bench_r2_a.json 2> $O/bench_r2_a.err; echo "bench rc=$?"; cut -c1-300 $O/bench_r2_a.json
for p in 4 8; do python bench.py --parties $p --steps 3 --warmup 3 > $O/bench_r2_a_${p}party.json 2> $O/bench_r2_a_${p}party.err; cut -c1-200 $O/bench_r2_a_${p}party.json; done
