#!/bin/bash
# round 2, final build on 2 GPUs: the driver's own command line (torchrun, default workload with extras: weak + strong scaling, in-process context)
mkdir -p gpurun_out
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/bench_n8_final.json 2> gpurun_out/bench_n8_final.err
tail -4 gpurun_out/bench_n8_final.err
python - <<'PY'
import json
for l in open("gpurun_out/bench_n8_final.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(round(d["value"]), "e2e", round(d["e2e"]["value"]), d["n_gpus"])
        ex = d.get("extras", {})
        for k, v in ex.get("strong_scaling", {}).items():
            print("strong", k, v if not isinstance(v, dict) else {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items()})
        for k, v in ex.get("in_process", {}).items():
            print("in_process", k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items()} if isinstance(v, dict) else v)
PY
