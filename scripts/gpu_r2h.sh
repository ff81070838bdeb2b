#!/bin/bash
# round 2, visit h: concurrent decodes / side-stream commitment check: parity tests, latency breakdown, verifier + proof bench lines
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/pytest_gpu_r2h.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2h.log
timeout 600 python scripts/latency_breakdown.py > gpurun_out/latency_breakdown_r2h.txt 2>&1
head -9 gpurun_out/latency_breakdown_r2h.txt
for w in verify_cells verify_blob_batch blob_proof verify_cells_one_batch; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-extras > gpurun_out/bench_${w}_r2h.json 2> gpurun_out/bench_${w}_r2h.err
  tail -2 gpurun_out/bench_${w}_r2h.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_r2h.json')); print('$w', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), d['kernel_ms_per_step'], d['oracle_check'])"
done
