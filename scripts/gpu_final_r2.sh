#!/bin/bash
# round 2, final visit (1 GPU): parity suite, the bench line of every workload (CPU-port baseline included), the reference arm,
# launch lists of the default bench and of the cell verifier.  Outputs under gpurun_out/, copied to profiles/r02_* afterwards.
mkdir -p gpurun_out
TAG=${1:-r2final}
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/pytest_gpu_$TAG.log 2>&1
tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_default_$TAG.json 2> gpurun_out/bench_default_$TAG.err
for w in commit blob_proof verify_blob_batch recover verify_cells verify_cells_one_batch; do
  timeout 600 python bench.py --workload $w > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_default_bench_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_default_$TAG.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_verify_cells_$TAG.csv \
    python bench.py --workload verify_cells --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_verify_cells_$TAG.log 2>&1
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*_$TAG.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d.get("value", 0)), d.get("unit"), "e2e", round(d.get("e2e", {}).get("value", 0)), "ms", round(d.get("ms_per_step", 0), 2), d.get("oracle_check"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
