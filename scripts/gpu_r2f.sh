#!/bin/bash
# round 2, sixth visit: both pairing forms + the overlapped FK20 chains: parity tests, pairing sweep, bench lines of the three BASELINE metrics
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/pytest_gpu_r2f.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2f.log
timeout 600 python scripts/pairing_sweep.py > gpurun_out/pairing_sweep.log 2>&1
grep "^n =" gpurun_out/pairing_sweep.log
for w in cells_proofs recover verify_cells verify_blob_batch; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-extras > gpurun_out/bench_${w}_r2f.json 2> gpurun_out/bench_${w}_r2f.err
  tail -2 gpurun_out/bench_${w}_r2f.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_r2f.json')); print('$w', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), d['kernel_ms_per_step'], d['oracle_check'])"
done
