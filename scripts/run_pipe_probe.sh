#!/bin/bash
# builds (if nvcc is there) and runs the pipe microbenchmarks; JSON -> gpurun_out/pipe_probe.json
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out go-eth-kzg_b200/csrc/probe/bin
B=go-eth-kzg_b200/csrc/probe/bin/pipe_probe
if [ ! -x $B ] || [ go-eth-kzg_b200/csrc/probe/pipe_probe.cu -nt $B ] || [ go-eth-kzg_b200/csrc/fp64.cuh -nt $B ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -o $B go-eth-kzg_b200/csrc/probe/pipe_probe.cu
fi
$B | tee gpurun_out/pipe_probe.json
