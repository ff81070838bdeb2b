#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/decode_sweep.py > gpurun_out/decode_sweep.log 2>&1
cat gpurun_out/decode_sweep.log | tail -13
bash scripts/gpu_sanitize_r2.sh
