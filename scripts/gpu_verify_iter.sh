#!/bin/bash
# quick iteration on the verifier paths: parity tests, the two verifier bench lines, launch lists
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_verify.py tests/test_gpu_fullsize.py -x -q ) > gpurun_out/pytest_verify.log 2>&1
tail -5 gpurun_out/pytest_verify.log
for w in verify_cells verify_blob_batch; do
  python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  tail -2 gpurun_out/bench_$w.err
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$w.csv \
    python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$w.log 2>&1
done
cat gpurun_out/bench_verify_*.json | cut -c1-300
