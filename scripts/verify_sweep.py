"""Verifier tunables on ONE context (kernel-class ms from the library's CUDA events, min of 3, self-check each time):
large_window / large_item for the one-verdict cell workload, optimistic on/off for 4096 x 128-cell verdicts,
rlc_item for the EIP-4844 batch verdict.  Run on a GPU box."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch, kzgb200
from bench import make_work

def kms(ctx, w, reps=3):
    w.step(False)          # (the self-check also looks at the host-pointer path's results)
    w.step(True)
    best = None
    for _ in range(reps):
        w.step(True)
        k = {a: round(b, 3) for a, b in ctx.last_kernel_ms().items() if b}
        k["total"] = round(ctx.last_device_ms(), 3)
        if best is None or k["total"] < best["total"]:
            best = k
    return best

def tun(ctx, name, v):
    assert ctx.L.kzgb200_dbg_set_tunable(name, v) == 0, name

out = {}
ctx = kzgb200.Context(commit_window=8, fk20_window=8)
w = make_work(ctx, "verify_cells_one_batch", 4096, 0, torch, np, 0)
for lw, li in ((8, 0), (4, 128)):
    tun(ctx, b"large_window", lw); tun(ctx, b"large_item", li)
    r = kms(ctx, w); r["self_check"] = bool(w.self_check())
    out["one_batch large_window=%d large_item=%d" % (lw, li)] = r
    print("one_batch large_window", lw, "large_item", li, r, flush=True)
tun(ctx, b"large_window", 4); tun(ctx, b"large_item", 0)
del w; torch.cuda.empty_cache()
w = make_work(ctx, "verify_cells", 4096, 0, torch, np, 0)
for opt, ov in ((0, 0), (0, 1), (1, 0), (1, 1), (1, 0), (1, 1)):
    tun(ctx, b"optimistic", opt); tun(ctx, b"verify_overlap", ov)
    r = kms(ctx, w); r["self_check"] = bool(w.self_check())
    out.setdefault("verify_cells optimistic=%d verify_overlap=%d" % (opt, ov), []).append(r)
    print("verify_cells optimistic", opt, "verify_overlap", ov, r, flush=True)
tun(ctx, b"optimistic", 1); tun(ctx, b"verify_overlap", 1)
del w; torch.cuda.empty_cache()
w = make_work(ctx, "verify_blob_batch", 4096, 0, torch, np, 0)
for it in (128, 64, 32, 16):
    tun(ctx, b"rlc_item", it)
    r = kms(ctx, w); r["self_check"] = bool(w.self_check())
    out["verify_blob_batch rlc_item=%d" % it] = r
    print("verify_blob_batch rlc_item", it, r, flush=True)
tun(ctx, b"rlc_item", 0)
ctx.close()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "verify_sweep.json"), "w"), indent=1)
