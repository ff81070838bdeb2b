"""Does cutting ONE big host-buffer call into k concurrent ranges on the SAME GPU (k lanes, own streams + scratch) shorten it?
The context's device list names GPU 0 k times, so the library's own sharding path (kzgb200_api.cu: run_ranges) does the cutting:
latency-bound kernels of one range (pairing, hashing, interpolation) run beside the pipe-bound ones (decode, bucket MSM) of another.
Wall-clock of the C-ABI call on pinned host buffers, min of `reps`.  Run on a GPU box."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch, kzgb200
from bench import make_work

out = {}
for wl, B, fw in (("verify_cells", 4096, 8), ("verify_blob_batch", 4096, 8), ("cells_proofs", 1024, 13)):
    for k in (1, 2, 3, 4):
        if wl == "cells_proofs" and k > 2:
            continue          # k replicas of the FK20 table
        ctx = kzgb200.Context(commit_window=8, fk20_window=fw, devices=[0] * k)
        w = make_work(ctx, wl, B, 0, torch, np, 0)
        w.step(False); w.step(False)
        best = 1e9
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter(); w.step(False); best = min(best, time.perf_counter() - t0)
        ok = bool(w.host_check()) if getattr(w, "host_check", None) else None
        out["%s k=%d" % (wl, k)] = {"e2e_ms": round(best * 1e3, 2), "units_per_s": round(w.units / best), "host_check": ok}
        print(wl, "k", k, out["%s k=%d" % (wl, k)], flush=True)
        w.release(); del w; ctx.close(); del ctx; torch.cuda.empty_cache()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "lane_split_probe.json"), "w"), indent=1)
