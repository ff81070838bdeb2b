"""Host-buffer (e2e) EIP-4844 batch verdict: wall-clock against the library's device-side class times. Run on a GPU box."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch, kzgb200, ctypes
from bench import make_work
ctx = kzgb200.Context(commit_window=8, fk20_window=8)
w = make_work(ctx, "verify_blob_batch", 4096, 0, torch, np, 0)
for on_dev in (True, False):
    w.step(on_dev); w.step(on_dev)
    best = None
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter(); w.step(on_dev); dt = (time.perf_counter() - t0) * 1e3
        k = {a: round(b, 2) for a, b in ctx.last_kernel_ms().items() if b}
        if best is None or dt < best[0]:
            best = (round(dt, 2), round(ctx.last_device_ms(), 2), k)
    print("verify_blob_batch", "device buffers" if on_dev else "host buffers", "wall", best[0], "device span", best[1], best[2], flush=True)
# raw H2D of the same bytes from the same pinned buffer, one copy
h = w.keep[0]; d = torch.empty_like(h, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("plain H2D of the blobs: %.2f ms (%.1f GB/s)" % (dt * 1e3, h.numel() / dt / 1e9))
ctx.close()
