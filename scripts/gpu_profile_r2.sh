#!/bin/bash
# round 2: ncu --set full captures of the dominant kernels (raw CSV only: the .ncu-rep files exceed gpurun's 64 MiB return limit)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none"
# cells + proofs, 1024 blobs, fk20 window 14: skip the first warm-up step (1 + 8 x 14 launches), then k_msm_fixed and the first three G1 FFT stage launches
timeout 600 $NCU -k regex:"k_msm_fixed|k_g1fft_stage" -s 113 -c 4 -o /tmp/r02_ncu_cells_proofs -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_cells_proofs.log 2>&1
timeout 600 $NCU -k regex:"k_g1_check|k_vmsm_buckets|k_pairing_lanes" -s 6 -c 4 -o /tmp/r02_ncu_verify_cells -f \
  python bench.py --workload verify_cells --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_verify_cells.log 2>&1
for r in cells_proofs verify_cells; do
  ncu -i /tmp/r02_ncu_$r.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$r.raw.csv 2>/dev/null
  python scripts/ncu_raw_summary.py gpurun_out/r02_ncu_$r.raw.csv > gpurun_out/r02_ncu_$r.summary.md 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_default_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_default.log 2>&1
ls -la /tmp/*.ncu-rep gpurun_out/
