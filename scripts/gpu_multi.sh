#!/bin/bash
# usage: gpu_multi.sh N "<workloads>"   -- torchrun bench at N GPUs of one box
N=$1
mkdir -p gpurun_out
for w in $2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $w --no-cpu-baseline \
     > gpurun_out/scale_${w}_n$N.json 2> gpurun_out/scale_${w}_n$N.err
  tail -2 gpurun_out/scale_${w}_n$N.err
  cut -c1-260 gpurun_out/scale_${w}_n$N.json
done
