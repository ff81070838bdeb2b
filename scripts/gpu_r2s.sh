#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_cells.py tests/test_gpu_proofs_recovery.py -m gpu -x -q ) > gpurun_out/pytest_gpu_r2s.log 2>&1
tail -5 gpurun_out/pytest_gpu_r2s.log
timeout 600 python scripts/g1_midbatch_sweep.py > gpurun_out/g1_midbatch_sweep.log 2>&1
tail -10 gpurun_out/g1_midbatch_sweep.log
