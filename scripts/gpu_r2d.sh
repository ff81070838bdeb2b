#!/bin/bash
# round 2, fourth visit: parity tests (new single-verify prep), G1 FFT occupancy / split sweep, ncu of the G1 FFT stages at full occupancy, latency
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/pytest_gpu_r2d.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2d.log
timeout 900 python scripts/tunables_sweep.py > gpurun_out/tunables_sweep_r2d.log 2>&1
grep -v "^\[" gpurun_out/tunables_sweep_r2d.log | tail -12
timeout 600 python scripts/latency_breakdown.py > gpurun_out/latency_breakdown_r2d.txt 2>&1
head -9 gpurun_out/latency_breakdown_r2d.txt
# one launch per stage over all 1024 blobs (split 1): the profiler then sees the stage kernel at its real occupancy
KZGB200_G1FFT_SPLIT=1 timeout 600 ncu --set full --clock-control none -k regex:"k_g1fft_stage" -s 45 -c 3 -o /tmp/r02d_g1fft -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02d_ncu_g1fft.log 2>&1
ncu -i /tmp/r02d_g1fft.ncu-rep --page raw --csv > gpurun_out/r02d_ncu_g1fft.raw.csv 2>/dev/null
python scripts/ncu_raw_summary.py gpurun_out/r02d_ncu_g1fft.raw.csv > gpurun_out/r02d_ncu_g1fft.summary.md 2>&1
head -24 gpurun_out/r02d_ncu_g1fft.summary.md | cut -c1-200
