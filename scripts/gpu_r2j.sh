#!/bin/bash
# round 2, visit j (2 GPUs): two-warp SHA-256 (parity + timing against the one-thread form), then the default bench under torchrun
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/pytest_gpu_r2j.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2j.log
timeout 600 python scripts/latency_breakdown.py > gpurun_out/latency_breakdown_r2j.txt 2>&1
head -9 gpurun_out/latency_breakdown_r2j.txt
for w in verify_blob_batch blob_proof; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-extras > gpurun_out/bench_${w}_r2j.json 2> gpurun_out/bench_${w}_r2j.err
  tail -2 gpurun_out/bench_${w}_r2j.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_r2j.json')); print('$w', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), d['kernel_ms_per_step'], d['oracle_check'])"
done
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/bench_n2_r2j.json 2> gpurun_out/bench_n2_r2j.err
tail -4 gpurun_out/bench_n2_r2j.err
python - <<'PY'
import json
for l in open("gpurun_out/bench_n2_r2j.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(round(d["value"]), "e2e", round(d["e2e"]["value"]), d["n_gpus"])
        ip = d.get("extras", {}).get("in_process", {})
        for k, v in ip.items():
            print(k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items()} if isinstance(v, dict) else v)
PY
