#!/bin/bash
# round 2, visit k: paired-squaring sweep (G1 FFT stage, decode) and same-GPU lane split probe
mkdir -p gpurun_out
timeout 900 python scripts/dual_sweep.py > gpurun_out/dual_sweep.log 2>&1
cat gpurun_out/dual_sweep.log | tail -12
timeout 900 python scripts/lane_split_probe.py > gpurun_out/lane_split_probe.log 2>&1
cat gpurun_out/lane_split_probe.log | tail -12
