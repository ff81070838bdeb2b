"""kernel-class breakdown (CUDA events) of one-blob calls; run on a GPU box"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kzgb200, oracle_lib
ctx = kzgb200.Context(commit_window=13, fk20_window=13)
blob = oracle_lib.rand_blob(123 << 20)
st, cm = ctx.blob_to_kzg_commitment(blob)
z = (5).to_bytes(32, "big")
st, pz, y = ctx.compute_kzg_proof(blob, z)
st, bp = ctx.compute_blob_kzg_proof(blob, cm)
st, cells, proofs = ctx.compute_cells_and_kzg_proofs(blob)
cl = [cells[2048 * i:2048 * i + 2048] for i in range(128)]; pl = [proofs[48 * i:48 * i + 48] for i in range(128)]
calls = {"commit": lambda: ctx.blob_to_kzg_commitment(blob), "kzg_proof": lambda: ctx.compute_kzg_proof(blob, z), "blob_proof": lambda: ctx.compute_blob_kzg_proof(blob, cm),
         "verify_kzg": lambda: ctx.verify_kzg_proof(cm, z, y, pz), "verify_blob": lambda: ctx.verify_blob_kzg_proof(blob, cm, bp),
         "verify_blob_batch1": lambda: ctx.verify_blob_kzg_proof_batch([blob], [cm], [bp]), "cells_proofs": lambda: ctx.compute_cells_and_kzg_proofs(blob),
         "recover": lambda: ctx.recover_cells_and_kzg_proofs(list(range(0, 128, 2)), cl[0::2]),
         "verify_cells128": lambda: ctx.verify_cell_kzg_proof_batch([cm] * 128, list(range(128)), cl, pl)}
for name, f in calls.items():
    for _ in range(3):
        f()
    print(name, "%.3f ms" % ctx.last_device_ms(), {k: round(v, 3) for k, v in ctx.last_kernel_ms().items() if v})
for n in (2, 4, 8, 16, 32, 64, 128):
    blobs = [blob] * n
    for _ in range(2):
        ctx.compute_cells_and_kzg_proofs_batch(blobs)
    print("cells_proofs n=%d" % n, "%.3f ms" % ctx.last_device_ms(), {k: round(v, 3) for k, v in ctx.last_kernel_ms().items() if v})
