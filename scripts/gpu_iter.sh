#!/bin/bash
# usage: gpu_iter.sh "<pytest args>" "<workloads>" ["<workloads to launch-list>"]
mkdir -p gpurun_out
( time timeout 900 python -m pytest $1 -x -q ) > gpurun_out/pytest_iter.log 2>&1
tail -4 gpurun_out/pytest_iter.log
for w in $2; do
  python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  tail -2 gpurun_out/bench_$w.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$w.json"))
print("$w", round(d["value"]), d["unit"], "ms", round(d["ms_per_step"],2), d.get("kernel_ms_per_step"), "e2e", round(d["e2e"]["value"]))
PY
done
for w in $3; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$w.csv \
    python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$w.log 2>&1
done
