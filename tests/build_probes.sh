#!/bin/sh
# Builds the stand-alone tcgen05 probes (test infrastructure) into tests/bin/ (git-ignored; travels with gpurun).
#   tc_probe  modeA modeB swapA swapB [N] [K]   correctness of the UMMA descriptors / operand tile layout vs FP64
#   tc_probe2                                    cycles per 3xTF32 MMA batch for N = 64 / 128 / 256
set -e
cd "$(dirname "$0")"
mkdir -p bin
for p in tc_probe tc_probe2 tc_probe3; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 --extended-lambda -o bin/$p $p.cu
done
