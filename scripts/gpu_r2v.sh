#!/bin/bash
# round 2, visit v: tail overlap of the cell verifier (two lanes per GPU, gated): parity suite, then the two verify_cells bench lines
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=3 ) > gpurun_out/pytest_gpu_r2v.log 2>&1
tail -6 gpurun_out/pytest_gpu_r2v.log
for w in verify_cells verify_cells_one_batch; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_${w}_r2v.json 2> gpurun_out/bench_${w}_r2v.err
  tail -2 gpurun_out/bench_${w}_r2v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_r2v.json')); print('$w', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['ms_per_step'],2), d['kernel_ms_per_step'], d['oracle_check'], (d.get('per_verdict_pass') or {}).get('value'))"
done
