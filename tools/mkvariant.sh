#!/bin/bash
# tools/mkvariant.sh <name> [nvcc -D flags...]  ->  build/lib_<name>.so  (kernel experiment for tools/ab.sh; prints the throughput kernel's resources)
name=$1; shift
mkdir -p build
cd torus-fhe_b200/csrc
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v "$@" -DMK_BUILD_ID="\"$name\"" -shared \
     -o ../../build/lib_$name.so mktfhe_b200.cu -lcudart -ldl 2>&1 | grep -A2 -E "blind_rotate_(fft_)?kernelILi2ELi2" | grep -v "^--" | tr '\n' ' ' | sed "s/^/$name: /"; echo
