#!/bin/bash
# BASELINE configs[3] / [4] and the other workloads on the FFT build (one GPU)
O=gpurun_out; mkdir -p $O
python bench.py --workload adder --steps 2 --warmup 1 > $O/bench_r2_q_adder.json 2> $O/bench_r2_q_adder.err; echo "adder rc=$?"; cut -c1-300 $O/bench_r2_q_adder.json
python bench.py --workload adder --resident --steps 2 --warmup 1 > $O/bench_r2_q_adder_resident.json 2> $O/bench_r2_q_adder_resident.err; echo "adder resident rc=$?"; cut -c1-300 $O/bench_r2_q_adder_resident.json
python bench.py --workload less --resident --steps 2 --warmup 1 > $O/bench_r2_q_less_resident.json 2> $O/bench_r2_q_less_resident.err; echo "less rc=$?"; cut -c1-300 $O/bench_r2_q_less_resident.json
python bench.py --workload conv --steps 1 --warmup 1 > $O/bench_r2_q_conv_1gpu.json 2> $O/bench_r2_q_conv_1gpu.err; echo "conv rc=$?"; cut -c1-300 $O/bench_r2_q_conv_1gpu.json
python bench.py --workload single --steps 3 --warmup 2 > $O/bench_r2_q_single.json 2> $O/bench_r2_q_single.err; echo "single rc=$?"; cut -c1-300 $O/bench_r2_q_single.json
python bench.py --workload ccs --steps 2 --warmup 1 > $O/bench_r2_q_ccs.json 2> $O/bench_r2_q_ccs.err; echo "ccs rc=$?"; cut -c1-300 $O/bench_r2_q_ccs.json
