"""k_pairing_lanes: latency form (32 lanes per check) against throughput form (8 lanes) over the batch size; n independent VerifyKZGProof
checks of the same valid item, time of the PAIRING kernel class from the library's CUDA events.  Run on a GPU box."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kzgb200, oracle_lib
ctx = kzgb200.Context(commit_window=10, fk20_window=8)
blob = oracle_lib.rand_blob(7 << 20)
st, cm = ctx.blob_to_kzg_commitment(blob)
z = (12345).to_bytes(32, "big")
st, pz, y = ctx.compute_kzg_proof(blob, z)
out = {}
for n in (1, 64, 256, 592, 1024, 1776, 2048, 3072, 4096):
    row = {}
    for lanes in (32, 8, 0):
        assert ctx.L.kzgb200_dbg_set_tunable(b"pairing_lanes", lanes) == 0
        best = None
        for _ in range(3):
            res = ctx.verify_kzg_proof_batch([cm] * n, [z] * n, [y] * n, [pz] * n)
            assert all(r == 0 for r in res), res[:4]
            t = ctx.last_kernel_ms()["pairing"]
            best = t if best is None else min(best, t)
        row["lanes%d" % lanes] = round(best, 3)
    # a wrong y must be rejected by both forms
    for lanes in (32, 8):
        ctx.L.kzgb200_dbg_set_tunable(b"pairing_lanes", lanes)
        bad = ctx.verify_kzg_proof_batch([cm] * min(n, 8), [z] * min(n, 8), [(5).to_bytes(32, "big")] * min(n, 8), [pz] * min(n, 8))
        assert all(r == 1 for r in bad), bad
    out[n] = row
    print("n =", n, row, flush=True)
ctx.L.kzgb200_dbg_set_tunable(b"pairing_lanes", 0)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "pairing_sweep.json"), "w"), indent=1)
