#!/bin/bash
# round 2, visit o: ncu --set full of the cell verifier's kernels in the final build (optimistic combined pass: decode with paired
# squarings, column bucket MSM on 4-bit windows, interpolation, Horner, column twiddles)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none"
timeout 900 $NCU -k regex:"k_g1_check|k_vmsm_buckets|k_cell_interp$|k_cell_columns_large|k_vmsm_combine|k_vmsm_bucket_reduce" -s 9 -c 9 -o /tmp/r02o_verify -f \
  python bench.py --workload verify_cells --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02o_ncu_verify.log 2>&1
ncu -i /tmp/r02o_verify.ncu-rep --page raw --csv > gpurun_out/r02o_ncu_verify.raw.csv 2>/dev/null
python scripts/ncu_raw_summary.py gpurun_out/r02o_ncu_verify.raw.csv > gpurun_out/r02o_ncu_verify.summary.md 2>&1
head -12 gpurun_out/r02o_ncu_verify.summary.md | cut -c1-400
tail -3 gpurun_out/r02o_ncu_verify.log
