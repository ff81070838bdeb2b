#!/bin/bash
# round 2, visit t: three-piece proof path, recover's cells D2H under FK20: parity suite + the two bench lines
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=3 ) > gpurun_out/pytest_gpu_r2t.log 2>&1
tail -5 gpurun_out/pytest_gpu_r2t.log
for w in recover blob_proof; do
  timeout 600 python bench.py --workload $w > gpurun_out/bench_${w}_r2t.json 2> gpurun_out/bench_${w}_r2t.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_r2t.json')); print('$w', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['ms_per_step'],2), d['oracle_check'])"
done
