#!/bin/bash
# round 2, visit m: large-verdict column MSM on 4-bit windows, optimistic combined pass, rlc_item sweep; whole GPU suite
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/pytest_gpu_r2m.log 2>&1
tail -12 gpurun_out/pytest_gpu_r2m.log
timeout 900 python scripts/verify_sweep.py > gpurun_out/verify_sweep.log 2>&1
tail -14 gpurun_out/verify_sweep.log
