#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 KZG engine (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cells_proofs|commit] [--blobs B]
    python bench.py --impl reference ...        # the CPU arm (oracle port on the host cores)

A "step" is one pass of the hot path over one batch of synthetic blobs:
  cells_proofs (default): Context.ComputeCellsAndKZGProofs on B = 1024 blobs per GPU
                          (BASELINE.json configs[2]); metric = blobs/s
  commit:                 Context.BlobToKZGCommitment on B = 4096 blobs per GPU (configs[1], first leg)
Blob b of rank r is the reference generator GetRandBlob(seed = (r*B + b) << 20)
(bench_test.go:17-46: scalar j = SHA-256(be_int64(seed + 32 j)) mod r).

value   : device-resident throughput (inputs in HBM before the timed region; C-ABI called on
          device pointers), whole job over all ranks, max-over-ranks time.
e2e     : same calls on PINNED HOST buffers: H2D of the blobs and D2H of cells+proofs+status are
          inside the timed region.
Multi-GPU: one process per GPU (torchrun), blobs sharded by rank, no data-path collective;
torch.distributed is used only for the barrier and the max-over-ranks of the step time.
"""
import argparse, ctypes, hashlib, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200"))

R_MOD = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
BLOB = 131072

# Work model of SURVEY.md section 8(d): IMAD per Fp mul = 588, Fr mul = 264
IMAD_FP_MUL = 588
CANONICAL_W = {"cells_proofs": 2605e6, "commit": 549e6}   # IMAD per blob, SURVEY 8(d)


def rand_blob(seed):
    out = bytearray(BLOB)
    sha = hashlib.sha256
    for j in range(4096):
        d = sha((seed + 32 * j).to_bytes(8, "big", signed=True)).digest()
        out[32 * j:32 * j + 32] = (int.from_bytes(d, "big") % R_MOD).to_bytes(32, "big")
    return bytes(out)


def make_blobs(first, count):
    """count distinct blobs; generated in parallel on the host cores"""
    seeds = [(first + b) << 20 for b in range(count)]
    try:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(os.cpu_count() or 1, 16)) as pool:
            return pool.map(rand_blob, seeds, chunksize=8)
    except Exception:
        return [rand_blob(s) for s in seeds]


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline(workload, seconds_budget=20.0):
    """the oracle port (reference's algorithmic structure, C++) on all host cores, bounded sample.
    Returns (cpu_baseline object, units in the sample, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    o = oracle_lib.get_oracle()
    cores = os.cpu_count() or 1
    L = oracle_lib.lib()
    note = "C++ restatement of the reference algorithm (oracle/), NOT gnark-crypto: no Go toolchain in this image"
    if workload in ("commit", "cells_proofs"):
        kind = 0 if workload == "commit" else 1
        per = 0.012 if kind == 0 else 0.45                       # rough single-thread seconds per blob
        n = max(cores, int(seconds_budget * cores / per / 2))
        n = min(n, 64 * cores)
        blobs = b"".join(make_blobs(1 << 30, min(n, 32)) * ((n + 31) // 32))[:n * BLOB]
        a = ctypes.create_string_buffer(n * (48 if kind == 0 else 262144))
        b = ctypes.create_string_buffer(n * 6144 if kind else 1)
        L.ko_parallel_blobs(o.ctx, kind, blobs, ctypes.c_size_t(min(n, cores)), cores, a, b)   # warm-up
        t = time.perf_counter()
        rc = L.ko_parallel_blobs(o.ctx, kind, blobs, ctypes.c_size_t(n), cores, a, b)
        dt = time.perf_counter() - t
        assert rc == 0
        return {"value": n / dt, "unit": "blobs/s", "cores": cores, "kind": "port",
                "sample": f"{n} blobs of the same generator, {cores} host threads, blob-parallel, one pass ({dt:.1f} s); " + note}, n, dt
    # the other workloads: inputs are prepared with the oracle itself (untimed), then the timed call runs once per blob on a
    # pool of `cores` threads (ctypes releases the GIL), repeated until the budget is used
    from concurrent.futures import ThreadPoolExecutor
    nb = 2 * cores
    blobs = make_blobs(1 << 30, nb)
    pool = ThreadPoolExecutor(cores)
    cms = list(pool.map(lambda b: o.blob_to_kzg_commitment(b)[1], blobs))
    unit, per_call_units = "blobs/s", 1
    if workload == "blob_proof":
        call = lambda i: o.compute_blob_kzg_proof(blobs[i], cms[i])[0]
        what = "ComputeBlobKZGProof per blob"
    elif workload == "verify_blob_batch":
        pfs = list(pool.map(lambda i: o.compute_blob_kzg_proof(blobs[i], cms[i])[1], range(nb)))
        half = nb // cores
        call = lambda i: o.verify_blob_kzg_proof_batch(blobs[i * half:(i + 1) * half], cms[i * half:(i + 1) * half], pfs[i * half:(i + 1) * half]) if i < cores else 0
        per_call_units = half
        what = f"VerifyBlobKZGProofBatch over {half} blobs per thread (the reference's batch loop is sequential, verify.go:102-140)"
    else:
        full = list(pool.map(lambda b: o.compute_cells_and_kzg_proofs(b), blobs))
        cells = [[f[1][2048 * i:2048 * i + 2048] for i in range(128)] for f in full]
        proofs = [[f[2][48 * i:48 * i + 48] for i in range(128)] for f in full]
        if workload == "recover":
            import numpy as np
            ids = [sorted(np.random.default_rng(i).choice(128, 64, replace=False).tolist()) for i in range(nb)]
            call = lambda i: o.recover_cells_and_kzg_proofs(ids[i], [cells[i][k] for k in ids[i]])[0]
            what = "RecoverCellsAndComputeKZGProofs per blob, 64 random cells"
        else:
            call = lambda i: o.verify_cell_kzg_proof_batch([cms[i]] * 128, list(range(128)), cells[i], proofs[i])
            unit, per_call_units = "cells/s", 128
            what = "VerifyCellKZGProofBatch, one 128-cell verdict per call"
    n_calls = cores if workload == "verify_blob_batch" else nb
    assert all(st == 0 for st in pool.map(call, range(n_calls)))           # warm-up + correctness of the sample
    t = time.perf_counter()
    done = 0
    while True:
        assert all(st == 0 for st in pool.map(call, range(n_calls)))
        done += n_calls * per_call_units
        dt = time.perf_counter() - t
        if dt > seconds_budget / 2:
            break
    pool.shutdown()
    return {"value": done / dt, "unit": unit, "cores": cores, "kind": "port",
            "sample": f"{what}; {n_calls} calls per pass on {cores} host threads, {done} units in {dt:.1f} s; " + note}, done, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    base, n, dt = cpu_baseline(wl, seconds_budget=max(5.0, 60.0 / max(1, args.steps + args.warmup)))
    # steps: repeat the bounded sample
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    line = {"impl": "reference", "metric": METRIC[wl], "value": base["value"], "unit": base["unit"], "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 limbs (Fp 381-bit / Fr 255-bit Montgomery)", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME[wl], "sample_units_per_step": n},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": base["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


METRIC = {"cells_proofs": "ComputeCellsAndKZGProofs blobs/s", "commit": "BlobToKZGCommitment blobs/s",
          "blob_proof": "ComputeBlobKZGProof blobs/s", "verify_blob_batch": "VerifyBlobKZGProofBatch blobs/s",
          "recover": "RecoverCellsAndComputeKZGProofs blobs/s", "verify_cells": "VerifyCellKZGProofBatch cells/s",
          "verify_cells_one_batch": "VerifyCellKZGProofBatch cells/s"}
WORKLOAD_NAME = {"cells_proofs": "EIP-7594 ComputeCellsAndKZGProofs (FK20, 128 cells x 64 Fr), 1024 random blobs per GPU",
                 "commit": "EIP-4844 BlobToKZGCommitment, 4096 random blobs per GPU",
                 "blob_proof": "EIP-4844 ComputeBlobKZGProof, 4096 random blobs per GPU",
                 "verify_blob_batch": "EIP-4844 VerifyBlobKZGProofBatch, one RLC verdict over 4096 blobs per GPU",
                 "recover": "EIP-7594 RecoverCellsAndComputeKZGProofs, 64 of 128 cells (random pattern), 1024 blobs per GPU",
                 "verify_cells": "EIP-7594 VerifyCellKZGProofBatch, independent 128-cell batches, 128 x B cells per GPU",
                 "verify_cells_one_batch": "EIP-7594 VerifyCellKZGProofBatch, ONE verdict over all 128 x B cells per GPU (SURVEY 8d secondary shape)"}
DEFAULT_B = {"cells_proofs": 1024, "commit": 4096, "blob_proof": 4096, "verify_blob_batch": 4096, "recover": 1024, "verify_cells": 4096, "verify_cells_one_batch": 4096}
# canonical IMAD per unit
CANONICAL_W.update({"blob_proof": 560e6, "verify_blob_batch": 8.2e6, "recover": 2701e6, "verify_cells": 1.79e6, "verify_cells_one_batch": 1.27e6})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (default: 5; 25 / 10 for the short verifier steps so the clock sampler sees the timed region)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="cells_proofs", choices=sorted(METRIC))
    ap.add_argument("--blobs", type=int, default=0, help="blobs per GPU per step")
    ap.add_argument("--commit-window", type=int, default=15)
    ap.add_argument("--fk20-window", type=int, default=14)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.steps <= 0:
        args.steps = {"verify_blob_batch": 25, "verify_cells": 10, "verify_cells_one_batch": 10}.get(args.workload, 5)
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly ONE JSON line: everything libraries print to fd 1 (e.g. "NCCL version ..." with NCCL_DEBUG
    # set on the box) is sent to stderr, and the result line is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import kzgb200
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = args.workload
    B = args.blobs or DEFAULT_B[wl]
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dbg = kzgb200.Debug()
    # tables: only the one this workload needs gets the big window
    uses_commit = wl in ("commit", "blob_proof")
    uses_fk20 = wl in ("cells_proofs", "recover")
    cw = args.commit_window if uses_commit else 8
    fw = args.fk20_window if uses_fk20 else 8
    ctx = kzgb200.Context(device=local, commit_window=cw, fk20_window=fw)
    info0 = ctx.info()
    L = ctx.L
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    SZ = ctypes.c_size_t

    def pinned(nbytes, dtype=torch.uint8):
        return torch.empty(nbytes, dtype=dtype).pin_memory()

    blobs = make_blobs(rank * B, B)
    h_blobs = torch.frombuffer(bytearray(b"".join(blobs)), dtype=torch.uint8).pin_memory()
    del blobs
    d_blobs = h_blobs.cuda()
    d_st = torch.zeros(max(B, 1), dtype=torch.int32, device="cuda"); h_st = pinned(max(B, 1), torch.int32)
    units_per_step = B
    check = lambda: None

    if wl == "commit":
        d_o = torch.empty(48 * B, dtype=torch.uint8, device="cuda"); h_o = pinned(48 * B)
        def step(dev):
            i, o, s = (d_blobs, d_o, d_st) if dev else (h_blobs, h_o, h_st)
            ctx._check(L.kzgb200_blob_to_kzg_commitment(ctx.ctx, P(i), SZ(B), P(o), P(s)))
        h2d, d2h = B * BLOB, (48 + 4) * B
        same = lambda: bytes(h_o[:480].numpy()) == bytes(d_o[:480].cpu().numpy())
    elif wl == "blob_proof":
        d_c = torch.empty(48 * B, dtype=torch.uint8, device="cuda")
        ctx._check(L.kzgb200_blob_to_kzg_commitment(ctx.ctx, P(d_blobs), SZ(B), P(d_c), P(d_st)))
        h_c = d_c.cpu().pin_memory()
        d_o = torch.empty(48 * B, dtype=torch.uint8, device="cuda"); h_o = pinned(48 * B)
        def step(dev):
            i, c_, o, s = (d_blobs, d_c, d_o, d_st) if dev else (h_blobs, h_c, h_o, h_st)
            ctx._check(L.kzgb200_compute_blob_kzg_proof(ctx.ctx, P(i), P(c_), SZ(B), P(o), P(s)))
        h2d, d2h = B * (BLOB + 48), (48 + 4) * B
        same = lambda: bytes(h_o[:480].numpy()) == bytes(d_o[:480].cpu().numpy())
    elif wl == "verify_blob_batch":
        d_c = torch.empty(48 * B, dtype=torch.uint8, device="cuda"); d_p = torch.empty(48 * B, dtype=torch.uint8, device="cuda")
        ctx._check(L.kzgb200_blob_to_kzg_commitment(ctx.ctx, P(d_blobs), SZ(B), P(d_c), P(d_st)))
        ctx._check(L.kzgb200_compute_blob_kzg_proof(ctx.ctx, P(d_blobs), P(d_c), SZ(B), P(d_p), P(d_st)))
        h_c = d_c.cpu().pin_memory(); h_p = d_p.cpu().pin_memory()
        res = ctypes.c_int32(-1)
        def step(dev):
            i, c_, p_ = (d_blobs, d_c, d_p) if dev else (h_blobs, h_c, h_p)
            ctx._check(L.kzgb200_verify_blob_kzg_proof_batch(ctx.ctx, P(i), P(c_), P(p_), SZ(B), ctypes.byref(res)))
            assert res.value == 0, "valid batch rejected"
        # a corrupted proof must be rejected (Cfg2)
        bad = d_p.clone(); bad[17 * 48:18 * 48] = torch.frombuffer(bytearray(kzgb200.load_trusted_setup()[0][:48]), dtype=torch.uint8).cuda()
        ctx._check(L.kzgb200_verify_blob_kzg_proof_batch(ctx.ctx, P(d_blobs), P(d_c), P(bad), SZ(B), ctypes.byref(res)))
        assert res.value == 1, "corrupted batch accepted"
        h2d, d2h = B * (BLOB + 96), 4
        same = lambda: True
    else:
        # 7594 workloads need cells + proofs of the blobs
        d_cells = torch.empty(262144 * B, dtype=torch.uint8, device="cuda"); d_pr = torch.empty(6144 * B, dtype=torch.uint8, device="cuda")
        h_cells = pinned(262144 * B); h_pr = pinned(6144 * B)
        if wl == "cells_proofs":
            def step(dev):
                i, a, b, s = (d_blobs, d_cells, d_pr, d_st) if dev else (h_blobs, h_cells, h_pr, h_st)
                ctx._check(L.kzgb200_compute_cells_and_kzg_proofs(ctx.ctx, P(i), SZ(B), P(a), P(b), P(s)))
            h2d, d2h = B * BLOB, (262144 + 6144 + 4) * B
            same = lambda: bytes(h_pr[:480].numpy()) == bytes(d_pr[:480].cpu().numpy())
        else:
            # reference outputs from a small-window context would need a second table; compute with this ctx
            ctx._check(L.kzgb200_compute_cells_and_kzg_proofs(ctx.ctx, P(d_blobs), SZ(B), P(d_cells), P(d_pr), P(d_st)))
            assert int(d_st.abs().sum().item()) == 0
            if wl == "recover":
                ids = np.concatenate([np.sort(np.random.default_rng(rank * B + b).choice(128, 64, replace=False)) for b in range(B)]).astype(np.uint64)
                counts = np.full(B, 64, dtype=np.uint64)
                cells_np = d_cells.cpu().numpy().reshape(B, 128, 2048)
                sel = np.stack([cells_np[b, ids[64 * b:64 * b + 64].astype(np.int64)] for b in range(B)])      # [B,64,2048]
                h_in = torch.from_numpy(np.ascontiguousarray(sel.reshape(-1))).pin_memory(); d_in = h_in.cuda()
                d_oc = torch.empty(262144 * B, dtype=torch.uint8, device="cuda"); d_op = torch.empty(6144 * B, dtype=torch.uint8, device="cuda")
                h_oc = pinned(262144 * B); h_op = pinned(6144 * B)
                idp = ids.ctypes.data_as(ctypes.c_void_p); cnp = counts.ctypes.data_as(ctypes.c_void_p)
                def step(dev):
                    i, a, b, s = (d_in, d_oc, d_op, d_st) if dev else (h_in, h_oc, h_op, h_st)
                    ctx._check(L.kzgb200_recover_cells_and_kzg_proofs(ctx.ctx, idp, cnp, P(i), SZ(B), P(a), P(b), P(s)))
                h2d, d2h = B * 64 * 2048, (262144 + 6144 + 4) * B
                same = lambda: bool(torch.equal(d_oc, d_cells)) and bool(torch.equal(d_op, d_pr))      # recovery == direct computation
            else:   # verify_cells: B independent 128-cell batches
                d_cm1 = torch.empty(48 * B, dtype=torch.uint8, device="cuda")
                ctx._check(L.kzgb200_blob_to_kzg_commitment(ctx.ctx, P(d_blobs), SZ(B), P(d_cm1), P(d_st)))
                N = 128 * B
                d_cm = d_cm1.view(B, 1, 48).expand(B, 128, 48).contiguous().view(-1)
                h_cm = d_cm.cpu().pin_memory(); h_cells.copy_(d_cells); h_pr.copy_(d_pr)
                one = wl == "verify_cells_one_batch"
                idx = np.tile(np.arange(128, dtype=np.uint64), B)
                offs = np.array([0, N], dtype=np.uint64) if one else (np.arange(B + 1, dtype=np.uint64) * 128)
                nv = 1 if one else B
                d_res = torch.empty(nv, dtype=torch.int32, device="cuda"); h_res = pinned(nv, torch.int32)
                ip = idx.ctypes.data_as(ctypes.c_void_p); op_ = offs.ctypes.data_as(ctypes.c_void_p)
                def step(dev):
                    cm, ce, pr, rs = (d_cm, d_cells, d_pr, d_res) if dev else (h_cm, h_cells, h_pr, h_res)
                    ctx._check(L.kzgb200_verify_cell_kzg_proof_batch(ctx.ctx, P(cm), ip, P(ce), P(pr), SZ(N), op_, SZ(nv), P(rs)))
                units_per_step = N
                h2d, d2h = N * (48 + 2048 + 48), 4 * nv
                def same():
                    ok = int(d_res.abs().sum().item()) == 0 and int(h_res.abs().sum().item()) == 0
                    badc = d_cells.clone(); badc[5 * 2048 + 40] ^= 1                                   # corrupt one cell of batch 0
                    ctx._check(L.kzgb200_verify_cell_kzg_proof_batch(ctx.ctx, P(d_cm), ip, P(badc), P(d_pr), SZ(N), op_, SZ(nv), P(d_res)))
                    r = d_res.cpu()
                    return ok and int(r[0]) == 1 and int(r[1:].abs().sum()) == 0

    # ---- device-resident timing ---------------------------------------------------------------
    for _ in range(args.warmup):
        step(True)
    sampler = ClockSampler(local); sampler.start()
    l0 = ctx.info()["kernel_launches"]
    kms = {}
    dev_ms = 0.0
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(True)
        dev_ms += ctx.last_device_ms()
        for k, v in ctx.last_kernel_ms().items():
            kms[k] = kms.get(k, 0.0) + v
    barrier()
    wall = time.perf_counter() - t0
    launches = ctx.info()["kernel_launches"] - l0
    clocks = sampler.stop()
    wall = max_over_ranks(wall)
    dev_ms = max_over_ranks(dev_ms)
    assert int(d_st.abs().sum().item()) == 0, "an item failed"
    # ---- end-to-end timing (pinned host buffers through the same C ABI) -----------------------
    step(False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(False)
    barrier()
    e2e_wall = max_over_ranks(time.perf_counter() - t0)
    assert same(), "self-check failed (host path vs device path / accept-reject)"
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    unit = "cells/s" if wl.startswith("verify_cells") else "blobs/s"
    total_units = units_per_step * world * args.steps
    value = total_units / wall
    tab = info0
    peaks = {m: dbg.imad_peak(local, i) for i, m in enumerate(("mad_lo", "mad_hi", "mad_wide"))}
    # one IMAD.WIDE retires a full 32x32->64 product = 2 IMAD-equivalents
    peak = max(peaks["mad_lo"], 2 * peaks["mad_wide"])
    peak_src = "measured live (kzgb200_bench_imad): mad.lo %.2f, mad.hi %.2f, mad.wide %.2f T lane-ops/s; peak = max(lo, 2 x wide); " \
               "the wide-multiply (fmaheavy) rate 2 x %.2f = %.2f T is what bounds multi-limb products" \
               % (peaks["mad_lo"] / 1e12, peaks["mad_hi"] / 1e12, peaks["mad_wide"] / 1e12, peaks["mad_wide"] / 1e12, 2 * peaks["mad_wide"] / 1e12)
    roof = {"bound": "int32-imad", "peak": peak / 1e12, "unit": "TIMAD/s", "peak_source": peak_src, "traffic": None,
            "whole_step_canonical": {"W_imad_per_unit": CANONICAL_W[wl], "achieved": value / world * CANONICAL_W[wl] / 1e12,
                                     "frac": value / world * CANONICAL_W[wl] / peak}}
    cfg = {"workload": WORKLOAD_NAME[wl], "blobs_per_gpu_per_step": B,
           "l2_policy": "inputs larger than L2 (%.0f MB of input per step + multi-GB digit table gathered at random)" % (h2d / 1e6),
           "parallelism": "blob-sharded, %d process(es), no collective" % world}
    if uses_commit or uses_fk20:
        c_used, W_used = (tab["commit_window"], tab["commit_windows_per_scalar"]) if uses_commit else (tab["fk20_window"], tab["fk20_windows_per_scalar"])
        npts = 4096 if uses_commit else 8192
        # dominant kernel k_msm_fixed; executed work per blob: npts*W mixed additions of 10 Fp muls
        # a mixed addition is 8 Fp mul (588 IMAD each) + 2 Fp sqr; the dedicated squaring executes
        # 222 wide products + 12 lo = 456 IMAD-equivalents, so executed work is 5616 per addition
        # (SURVEY's nominal 5880 counts a squaring as a multiplication)
        IMAD_MIXED_ADD = 8 * IMAD_FP_MUL + 2 * 456
        msm_imad = npts * W_used * IMAD_MIXED_ADD
        msm_ms = kms.get("msm", 0.0) / args.steps
        achieved = msm_imad * B / (msm_ms * 1e-3) if msm_ms else None
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        roof.update({"kernel": "k_msm_fixed", "achieved": achieved / 1e12 if achieved else None, "frac": achieved / peak if achieved else None,
                     "frac_of_wide_multiply_rate": achieved / (2 * peaks["mad_wide"]) if achieved else None,
                     "work_model": "executed: %d pts x %d windows x (8 mul x 588 + 2 sqr x 456) IMAD = %.0f M IMAD/blob in this kernel (nominal at 588 per sqr: %.0f M)" % (npts, W_used, msm_imad / 1e6, npts * W_used * 10 * IMAD_FP_MUL / 1e6),
                     "hbm_table_gather": {"achieved": npts * W_used * 96.0 * B / (msm_ms * 1e-3) / 1e9 if msm_ms else None, "unit": "GB/s", "peak": hbm_peak}})
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get("cells_proofs" if uses_fk20 else "commit")
            if tr and tr["window_bits"] == c_used:
                roof["traffic"] = tr["bytes_per_launch"]
                roof["traffic_note"] = "ncu dram bytes per 1024-blob launch of k_msm_fixed (profiles/r01_ncu_full_headline.md); algorithmic gather bytes %d" % tr["algorithmic_gather_bytes_per_launch"]
        except Exception:
            pass
        cfg.update({"window_bits": c_used, "windows_per_scalar": W_used, "table_bytes": tab["commit_table_bytes"] if uses_commit else tab["fk20_table_bytes"]})
    else:
        # verifier workloads: executed-work model of the two throughput kernels
        #   k_g1_check per point: sqrt by 4-bit windows (384 sqr + 97 mul) + on-curve test (3) + subgroup test
        #     (126 Jacobian doublings of 2 mul + 5 sqr, 10 additions of 11 mul + 3 sqr, 2 x cached Z^2, Z^3)
        #   k_vmsm_buckets per point: 96 windows x 15/16 non-zero digits x mixed addition (8 mul + 2 sqr)
        SQR = 456
        DECODE = (384 * SQR + 100 * IMAD_FP_MUL) + 126 * (2 * IMAD_FP_MUL + 5 * SQR) + 10 * (11 * IMAD_FP_MUL + 3 * SQR) + 4 * IMAD_FP_MUL
        VMSM = 96 * 15 / 16 * (8 * IMAD_FP_MUL + 2 * SQR)
        pts_decode = {"verify_cells": units_per_step + B, "verify_cells_one_batch": units_per_step + B, "verify_blob_batch": 2 * B}[wl]      # proofs + unique commitments / proofs + commitments
        # bucket-MSM points in units of 96 four-bit windows: commitments of the 4844 batch carry 32 of 96; the one-verdict cell path uses 16 eight-bit windows
        pts_vmsm = {"verify_cells": units_per_step, "verify_cells_one_batch": units_per_step * (16 / (96 * 15 / 16)), "verify_blob_batch": 2 * B * (64 / 96)}[wl]
        per = {k: v / args.steps for k, v in kms.items() if v}
        models = {"decode": ("k_g1_check", pts_decode * DECODE), "vmsm": ("k_vmsm_buckets (+ reduce, combine)", pts_vmsm * VMSM)}
        top = max((k for k in per if k in models), key=lambda k: per[k], default=None)
        roof["kernel_classes"] = {k: {"kernel": models[k][0], "ms": per[k], "executed_imad": models[k][1],
                                      "achieved": models[k][1] / (per[k] * 1e-3) / 1e12, "frac": models[k][1] / (per[k] * 1e-3) / peak,
                                      "frac_of_wide_multiply_rate": models[k][1] / (per[k] * 1e-3) / (2 * peaks["mad_wide"])}
                                  for k in models if k in per}
        if top:
            t = roof["kernel_classes"][top]
            roof.update({"kernel": t["kernel"], "achieved": t["achieved"], "frac": t["frac"], "frac_of_wide_multiply_rate": t["frac_of_wide_multiply_rate"],
                         "work_model": "executed IMAD of the dominant kernel class '%s' (%.1f of %.1f ms device time): %.0f k IMAD per decoded point, %.0f k per bucket-MSM point; "
                                       "latency-bound tails (hashing, single pairing) are listed in kernel_ms_per_step" % (top, per[top], dev_ms / args.steps, DECODE / 1e3, VMSM / 1e3)})
        cfg["l2_policy"] = "inputs larger than L2 (%.0f MB of input per step)" % (h2d / 1e6)
    line = {
        "metric": METRIC[wl], "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (Fp 381-bit / Fr 255-bit Montgomery, integer)", "data": "synthetic", "config": cfg,
        "device_ms_per_step": dev_ms / args.steps,
        "kernel_ms_per_step": {k: v / args.steps for k, v in kms.items() if v},
        "gpu_launches": launches, "clocks": clocks,
        "e2e": {"value": total_units / e2e_wall, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "roofline": roof,
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(wl)[0]
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
