#!/bin/bash
# one GPU box visit: parity tests, the bench lines of every workload (with the CPU-port baseline), launch lists of the verifier workloads
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
python bench.py > gpurun_out/bench_cells_proofs.json 2> gpurun_out/bench_cells_proofs.err
for w in commit blob_proof verify_blob_batch recover verify_cells verify_cells_one_batch; do
  python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
if [ "$1" == "launches" ]; then
  for w in verify_blob_batch verify_cells; do
    ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$w.csv \
      python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$w.log 2>&1
  done
fi
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_*.json | cut -c1-200
