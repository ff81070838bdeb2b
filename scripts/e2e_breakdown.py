"""Host-buffer (e2e) calls of the 4844 proving paths: wall-clock against the library's own device-side class times. Run on a GPU box."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch, kzgb200
from bench import make_work
ctx = kzgb200.Context(commit_window=15, fk20_window=8)
for wl, pieces in (("blob_proof", 2), ("blob_proof", 3), ("blob_proof", 4), ("commit", 2)):
    assert ctx.L.kzgb200_dbg_set_tunable(b"proof_pieces", pieces) == 0
    w = make_work(ctx, wl, 4096, 0, torch, np, 0)
    for on_dev in ((True, False) if pieces == 2 else (False,)):
        w.step(on_dev); w.step(on_dev)
        best = None
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter(); w.step(on_dev); dt = (time.perf_counter() - t0) * 1e3
            k = {a: round(b, 2) for a, b in ctx.last_kernel_ms().items() if b}
            if best is None or dt < best[0]:
                best = (round(dt, 2), round(ctx.last_device_ms(), 2), k)
        if not on_dev:
            assert w.oracle_check()          # first and last blob of the host-path output against the CPU oracle
        print(wl, "pieces", pieces, "device buffers" if on_dev else "host buffers", "wall", best[0], "device span", best[1], best[2], flush=True)
    del w; torch.cuda.empty_cache()
ctx.close()
