"""k_g1_check (decompression + subgroup test) compiled for 4 / 6 / 8 resident CTAs per SM (255 / 168 / 128 registers) with and without paired
squarings: cell verifier, 4096 x 128 cells, kernel-class ms from the library's CUDA events (min of 3), checked every time.  Run on a GPU box."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch, kzgb200
from bench import make_work

def kms(ctx, w, reps=3):
    w.step(False); w.step(True)
    best = None
    for _ in range(reps):
        w.step(True)
        k = {a: round(b, 3) for a, b in ctx.last_kernel_ms().items() if b}
        k["total"] = round(ctx.last_device_ms(), 3)
        if best is None or k["total"] < best["total"]:
            best = k
    return best

out = {}
ctx = kzgb200.Context(commit_window=8, fk20_window=8)
w = make_work(ctx, "verify_cells", 4096, 0, torch, np, 0)
for ov in (0, 1):
    for dual in (0, 1):
        for minb in (4, 6, 8):
            for name, v in ((b"verify_overlap", ov), (b"decode_dual", dual), (b"decode_minb", minb)):
                assert ctx.L.kzgb200_dbg_set_tunable(name, v) == 0
            r = kms(ctx, w); r["self_check"] = bool(w.self_check()); r["oracle_check"] = bool(w.oracle_check())
            out["verify_overlap=%d decode_dual=%d decode_minb=%d" % (ov, dual, minb)] = r
            print("verify_overlap", ov, "decode_dual", dual, "decode_minb", minb, r, flush=True)
ctx.close()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "decode_sweep.json"), "w"), indent=1)
