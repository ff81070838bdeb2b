#!/bin/bash
# round 2, third visit: parity tests on the new defaults, tunable sweeps, ncu --set full of the new MSM / bucket kernels, bench
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/pytest_gpu_r2c.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2c.log
timeout 900 python scripts/tunables_sweep.py > gpurun_out/tunables_sweep.log 2>&1
grep -v "^\[" gpurun_out/tunables_sweep.log | tail -12
NCU="ncu --set full --clock-control none"
timeout 600 $NCU -k regex:"k_msm_fixed" -s 3 -c 1 -o /tmp/r02c_msm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02c_ncu_msm.log 2>&1
timeout 600 $NCU -k regex:"k_vmsm_buckets" -s 3 -c 1 -o /tmp/r02c_vmsm -f python bench.py --workload verify_cells --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_ncu_vmsm.log 2>&1
for r in msm vmsm; do
  ncu -i /tmp/r02c_$r.ncu-rep --page raw --csv > gpurun_out/r02c_ncu_$r.raw.csv 2>/dev/null
  python scripts/ncu_raw_summary.py gpurun_out/r02c_ncu_$r.raw.csv > gpurun_out/r02c_ncu_$r.summary.md 2>&1
done
for w in cells_proofs commit verify_cells; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-extras > gpurun_out/bench_${w}_r2c.json 2> gpurun_out/bench_${w}_r2c.err
  tail -2 gpurun_out/bench_${w}_r2c.err
  cut -c1-330 gpurun_out/bench_${w}_r2c.json
done
