#!/bin/bash
# ncu captures for profiles/: full sets of the dominant kernels + launch lists
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:"k_g1_check|k_vmsm_buckets|k_pairing_lanes" -c 4 -o gpurun_out/ncu_verify_cells -f \
  python bench.py --workload verify_cells --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_verify_cells.log 2>&1
$NCU -k regex:"k_msm_fixed|k_g1fft_stage" -s 24 -c 6 -o gpurun_out/ncu_cells_proofs -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_cells_proofs.log 2>&1
for r in verify_cells cells_proofs; do
  ncu -i gpurun_out/ncu_$r.ncu-rep --page raw --csv > gpurun_out/ncu_$r.raw.csv 2>/dev/null
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_default_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_default.log 2>&1
ls -la gpurun_out/*.ncu-rep
