"""k_msm_fixed variants (msm.cuh) on ONE context at the bench windows: kernel-class ms per variant for cells+proofs (1024 blobs)
and commitments (4096 blobs), outputs compared with variant 0 byte for byte.  Run on a GPU box."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, kzgb200, oracle_lib
sys.path.insert(0, ROOT)
from bench import make_blobs

def run(ctx, wl, B, variants, reps=3):
    L = ctx.L
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    blobs = torch.frombuffer(bytearray(b"".join(make_blobs(0, B))), dtype=torch.uint8).cuda()
    st = torch.zeros(B, dtype=torch.int32, device="cuda")
    res = {}
    ref = None
    for v in variants:
        assert L.kzgb200_dbg_set_tunable(b"msm_variant", v) == 0
        if wl == "cells_proofs":
            cells = torch.empty(B * 262144, dtype=torch.uint8, device="cuda"); out = torch.empty(B * 6144, dtype=torch.uint8, device="cuda")
            call = lambda: ctx._check(L.kzgb200_compute_cells_and_kzg_proofs(ctx.ctx, P(blobs), ctypes.c_size_t(B), P(cells), P(out), P(st)))
        else:
            out = torch.empty(B * 48, dtype=torch.uint8, device="cuda")
            call = lambda: ctx._check(L.kzgb200_blob_to_kzg_commitment(ctx.ctx, P(blobs), ctypes.c_size_t(B), P(out), P(st)))
        call()
        ms = []
        for _ in range(reps):
            call(); ms.append(ctx.last_kernel_ms().get("msm"))
        o = out.cpu()
        if ref is None: ref = o
        res[v] = {"msm_ms": min(ms), "total_ms": ctx.last_device_ms(), "same_as_v0": bool(torch.equal(o, ref)), "status_ok": int(st.abs().sum().item()) == 0}
        print(wl, "variant", v, res[v], flush=True)
    L.kzgb200_dbg_set_tunable(b"msm_variant", -1)
    return res

if __name__ == "__main__":
    variants = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else ",".join(str(i) for i in range(16))).split(",")]
    out = {}
    ctx = kzgb200.Context(commit_window=8, fk20_window=14)
    out["cells_proofs_1024_fk20_c14"] = run(ctx, "cells_proofs", 1024, variants)
    ctx.close(); torch.cuda.empty_cache()
    ctx = kzgb200.Context(commit_window=15, fk20_window=8)
    out["commit_4096_c15"] = run(ctx, "commit", 4096, variants)
    ctx.close()
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "msm_variants.json"), "w"), indent=1)
