#!/bin/bash
# Torus32 mode on the FFT channel: parity (single-key sets, CCS), then timing against the RNS kernels
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_tfhe1.py tests/test_ccs.py tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_fft_t32.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_fft_t32.log
for f in 1 0; do
  MKTFHE_B200_FFT=$f python bench.py --workload single --parties 80 --steps 3 --warmup 2 > $O/bench_t32_single80_fft$f.json 2> $O/bench_t32_single80_fft$f.err; echo "single80 fft=$f rc=$?"; cut -c1-160 $O/bench_t32_single80_fft$f.json
  MKTFHE_B200_FFT=$f python bench.py --workload ccs --steps 2 --warmup 1 > $O/bench_t32_ccs_fft$f.json 2> $O/bench_t32_ccs_fft$f.err; echo "ccs fft=$f rc=$?"; cut -c1-160 $O/bench_t32_ccs_fft$f.json
done
