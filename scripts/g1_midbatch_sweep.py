"""G1 transform of FK20 for medium batches: the two-level 16 x 8 form and the 4 x 4 x 4 x 2 form against the staged radix-2 form
(tunables g1_two_level_max, g1_chain4_max); kernel-class ms from the library's CUDA events, device-resident blobs, min of 3.  Run on a GPU box."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch, kzgb200
from bench import make_work

out = {}
ctx = kzgb200.Context(commit_window=8, fk20_window=13)
for n in (25, 32, 33, 64, 96, 128, 160, 192, 256):
    w = make_work(ctx, "cells_proofs", n, 0, torch, np, 0)
    row = {}
    for name, v, v4 in (("two_level", 1024, 0), ("chain4", 0, 1024), ("staged", 0, 0)):
        assert ctx.L.kzgb200_dbg_set_tunable(b"g1_two_level_max", v) == 0 and ctx.L.kzgb200_dbg_set_tunable(b"g1_chain4_max", v4) == 0
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
        row[name] = best
    ctx.L.kzgb200_dbg_set_tunable(b"g1_two_level_max", -1); ctx.L.kzgb200_dbg_set_tunable(b"g1_chain4_max", -1)
    out[n] = row
    print(n, "two_level", row["two_level"]["g1fft"], "chain4", row["chain4"]["g1fft"], "staged", row["staged"]["g1fft"], "totals", row["two_level"]["total"], row["chain4"]["total"], row["staged"]["total"],
          row["two_level"]["checks"], row["chain4"]["checks"], row["staged"]["checks"], flush=True)
    del w; torch.cuda.empty_cache()
ctx.close()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "g1_midbatch_sweep.json"), "w"), indent=1)
