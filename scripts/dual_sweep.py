"""Paired squarings (g1.cuh: fp_sqr2_ni, two independent Montgomery squarings in one out-of-line body) in the Jacobian doubling of
the G1 FFT stage kernel ("g1fft_dual") and of the subgroup test in the decode kernel ("decode_dual"): kernel-class times from the
library's CUDA events on ONE context per workload, min of 3, results checked by the work item's self-check and the CPU oracle.
Run on a GPU box."""
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
fw = int(sys.argv[1]) if len(sys.argv) > 1 else 14
ctx = kzgb200.Context(commit_window=8, fk20_window=fw)
w = make_work(ctx, "cells_proofs", 1024, 0, torch, np, 0)
for dual in (0, 1, 0, 1):
    assert ctx.L.kzgb200_dbg_set_tunable(b"g1fft_dual", dual) == 0
    r = kms(ctx, w); w.step(False); r["self_check"] = bool(w.self_check()); r["oracle_check"] = bool(w.oracle_check())
    out.setdefault("cells_proofs g1fft_dual=%d" % dual, []).append(r)
    print("cells_proofs g1fft_dual", dual, r, flush=True)
ctx.L.kzgb200_dbg_set_tunable(b"g1fft_dual", 0)
del w; torch.cuda.empty_cache()
w = make_work(ctx, "verify_cells", 4096, 0, torch, np, 0)
for dual in (0, 1, 0, 1):
    assert ctx.L.kzgb200_dbg_set_tunable(b"decode_dual", dual) == 0
    r = kms(ctx, w); w.step(False); r["self_check"] = bool(w.self_check()); r["oracle_check"] = bool(w.oracle_check())
    out.setdefault("verify_cells decode_dual=%d" % dual, []).append(r)
    print("verify_cells decode_dual", dual, r, flush=True)
ctx.L.kzgb200_dbg_set_tunable(b"decode_dual", 0)
ctx.close()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "dual_sweep.json"), "w"), indent=1)
