#!/bin/bash
# round 2, second visit: parity tests, k_msm_fixed variants, one-blob latency breakdown, the default bench line with every extra
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu_r2b.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2b.log
timeout 900 python scripts/msm_variants.py > gpurun_out/msm_variants.log 2>&1
cat gpurun_out/msm_variants.log | tail -20
timeout 600 python scripts/latency_breakdown.py > gpurun_out/latency_breakdown_r2b.txt 2>&1
cat gpurun_out/latency_breakdown_r2b.txt
timeout 900 python bench.py > gpurun_out/bench_default_r2b.json 2> gpurun_out/bench_default_r2b.err
tail -3 gpurun_out/bench_default_r2b.err
cut -c1-400 gpurun_out/bench_default_r2b.json
