"""Staged G1 transform at low occupancy: single-chain stage kernel against the dual-body one (tunable g1fft_dual), device ms of the
g1fft kernel class, min of 3, outputs checked.  Run on a GPU box."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch, kzgb200
from bench import make_work
out = {}
ctx = kzgb200.Context(commit_window=8, fk20_window=13)
for n in (33, 64, 128, 192, 256, 512):
    w = make_work(ctx, "cells_proofs", n, 0, torch, np, 0)
    row = {}
    for dual in (0, 1):
        assert ctx.L.kzgb200_dbg_set_tunable(b"g1fft_dual", dual) == 0
        w.step(True)
        best = None
        for _ in range(3):
            w.step(True)
            k = {a: round(b, 3) for a, b in ctx.last_kernel_ms().items() if b}
            k["total"] = round(ctx.last_device_ms(), 3)
            if best is None or k["total"] < best["total"]:
                best = k
        w.step(False)
        best["checks"] = bool(w.self_check()) and bool(w.oracle_check())
        row[dual] = best
    ctx.L.kzgb200_dbg_set_tunable(b"g1fft_dual", 0)
    out[n] = row
    print(n, "single", row[0]["g1fft"], row[0]["total"], "dual", row[1]["g1fft"], row[1]["total"], row[0]["checks"], row[1]["checks"], flush=True)
    del w; torch.cuda.empty_cache()
ctx.close()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "g1_dual_midbatch.json"), "w"), indent=1)
