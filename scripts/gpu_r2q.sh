#!/bin/bash
# round 2, visit q: row weights of small verdicts on the side chain: parity suite, latency breakdown, verify_cells bench line
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=3 ) > gpurun_out/pytest_gpu_r2q.log 2>&1
tail -6 gpurun_out/pytest_gpu_r2q.log
timeout 600 python scripts/latency_breakdown.py > gpurun_out/latency_breakdown_r2q.txt 2>&1
head -10 gpurun_out/latency_breakdown_r2q.txt
timeout 600 python bench.py --workload verify_cells --no-cpu-baseline > gpurun_out/bench_verify_cells_r2q.json 2> gpurun_out/bench_verify_cells_r2q.err
python -c "
import json; d=json.load(open('gpurun_out/bench_verify_cells_r2q.json')); print(round(d['value']), 'e2e', round(d['e2e']['value']), round(d['ms_per_step'],2), d['kernel_ms_per_step'], d['oracle_check']); print(d['per_verdict_pass'])"
