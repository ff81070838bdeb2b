#!/bin/bash
# round 2, visit i (2 GPUs): staggered-overlap experiment on GPU 0, then the default bench under torchrun with every extra
mkdir -p gpurun_out
timeout 600 python scripts/overlap_sweep.py > gpurun_out/overlap_sweep.log 2>&1
grep "fk20_overlap" gpurun_out/overlap_sweep.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/bench_n2_r2i.json 2> gpurun_out/bench_n2_r2i.err
tail -4 gpurun_out/bench_n2_r2i.err
python - <<'PY'
import json
for l in open("gpurun_out/bench_n2_r2i.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(round(d["value"]), "e2e", round(d["e2e"]["value"]), d["n_gpus"])
        e = d.get("extras", {})
        print(json.dumps(e.get("in_process"), indent=0)[:3000])
PY
