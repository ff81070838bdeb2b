#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=3 ) > gpurun_out/pytest_gpu_r2u.log 2>&1
tail -5 gpurun_out/pytest_gpu_r2u.log
timeout 600 python bench.py --workload verify_blob_batch > gpurun_out/bench_verify_blob_batch_r2u.json 2> gpurun_out/bench_verify_blob_batch_r2u.err
python -c "
import json; d=json.load(open('gpurun_out/bench_verify_blob_batch_r2u.json')); print(round(d['value']), 'e2e', round(d['e2e']['value']), round(d['ms_per_step'],2), d['oracle_check'])"
