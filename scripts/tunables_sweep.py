"""Run-time tunables (include/kzgb200_debug.h: kzgb200_dbg_set_tunable) measured on ONE context per workload:
fk20_lanes for cells+proofs (1024 blobs, fk20 window 14) and vmsm_policy for the cell verifier (4096 x 128 cells).
Kernel-class times from the library's CUDA events, results checked by the work item's own self-check.  Run on a GPU box."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch, kzgb200
from bench import make_work

def kms(ctx, w, reps=3):
    w.step(True)
    best = None
    for _ in range(reps):
        w.step(True)
        k = {a: round(b, 3) for a, b in ctx.last_kernel_ms().items() if b}
        k["total"] = round(ctx.last_device_ms(), 3)
        if best is None or k["total"] < best["total"]:
            best = k
    return best

out = {}
ctx = kzgb200.Context(commit_window=8, fk20_window=14)
w = make_work(ctx, "cells_proofs", 1024, 0, torch, np, 0)
w.step(False)
for minb, split in ((3, 8), (3, 4), (3, 2)):      # (the 4-CTA/SM build of the stage kernel, "g1fft_minb", was measured in round 2 and removed: 33.85 vs 33.71 ms)
    assert ctx.L.kzgb200_dbg_set_tunable(b"g1fft_split", split) == 0
    r = kms(ctx, w); r["self_check"] = bool(w.self_check())
    out["cells_proofs g1fft_minb=%d split=%d" % (minb, split)] = r
    print("cells_proofs g1fft_minb", minb, "split", split, r, flush=True)
ctx.L.kzgb200_dbg_set_tunable(b"g1fft_split", 0)
ctx.L.kzgb200_dbg_set_tunable(b"fk20_lanes", 0)
del w; torch.cuda.empty_cache()
w = make_work(ctx, "verify_cells", 4096, 0, torch, np, 0)
w.step(False)
for pol in (0, 1):
    assert ctx.L.kzgb200_dbg_set_tunable(b"vmsm_policy", pol) == 0
    r = kms(ctx, w); r["self_check"] = bool(w.self_check())
    out["verify_cells vmsm_policy=%d" % pol] = r
    print("verify_cells vmsm_policy", pol, r, flush=True)
ctx.L.kzgb200_dbg_set_tunable(b"vmsm_policy", 1)
del w; torch.cuda.empty_cache()
w = make_work(ctx, "verify_cells_one_batch", 4096, 0, torch, np, 0)
w.step(False)
for pol in (1,):
    assert ctx.L.kzgb200_dbg_set_tunable(b"vmsm_policy", pol) == 0
    r = kms(ctx, w); r["self_check"] = bool(w.self_check())
    out["verify_cells_one_batch vmsm_policy=%d" % pol] = r
    print("verify_cells_one_batch vmsm_policy", pol, r, flush=True)
ctx.L.kzgb200_dbg_set_tunable(b"vmsm_policy", 1)
ctx.close()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tunables_sweep.json"), "w"), indent=1)
