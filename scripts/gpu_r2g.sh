#!/bin/bash
# round 2: two GPUs of one box -- the multi-device tests, then the default bench under torchrun (strong scaling + in-process context extras)
mkdir -p gpurun_out
nvidia-smi -L | head -4
( time timeout 900 python -m pytest tests/test_gpu_multidev.py tests/test_gpu_threads.py -m gpu -x -q ) > gpurun_out/pytest_gpu_r2g.log 2>&1
tail -3 gpurun_out/pytest_gpu_r2g.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/bench_n2_r2g.json 2> gpurun_out/bench_n2_r2g.err
tail -5 gpurun_out/bench_n2_r2g.err
python - <<'PY'
import json
for l in open("gpurun_out/bench_n2_r2g.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(round(d["value"]), "e2e", round(d["e2e"]["value"]), d["n_gpus"])
        e = d.get("extras", {})
        print(json.dumps(e.get("strong_scaling"), indent=0)[:1500])
        print(json.dumps(e.get("in_process"), indent=0)[:3000])
PY
