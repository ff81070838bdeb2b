#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 KZG engine (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload W] [--blobs B]
    python bench.py --impl reference ...        # the CPU arm (oracle port on the host cores)
    python bench.py --workload latency          # single-call (n = 1) latency table of every Context method

A "step" is one pass of the hot path over one batch of synthetic blobs:
  cells_proofs (default): Context.ComputeCellsAndKZGProofs on B = 1024 blobs per GPU
                          (BASELINE.json configs[2]); metric = blobs/s
  commit / blob_proof / verify_blob_batch / recover / verify_cells / verify_cells_one_batch: the other configs.
Blob b of rank r is the reference generator GetRandBlob(seed = (r*B + b) << 20)
(bench_test.go:17-46: scalar j = SHA-256(be_int64(seed + 32 j)) mod r).

value   : device-resident throughput (inputs in HBM before the timed region; C-ABI called on
          device pointers), whole job over all ranks, max-over-ranks time.
e2e     : same calls on PINNED HOST buffers: H2D of the inputs and D2H of the outputs are inside the timed region.
After the timed region the outputs of the first and last blob are compared with the CPU oracle ("oracle_check").

The default line also carries `extras` (measured in the same run, fixed small step counts, stated per entry):
  coresident      : ONE context whose two window tables fit one GPU together (13/13): cells+proofs, commitments and
                    cell verification on it -- BASELINE.json's three metrics from a single drop-in context;
  latency_n1      : p50/p99 of one-blob calls through host pointers for every Context method, beside the oracle port;
  strong_scaling  : (torchrun, N > 1) a FIXED total batch (1024 blobs / 4096 blobs / 524 288 cells) sharded over the N ranks;
  in_process      : (torchrun, N > 1, rank 0 alone) ONE process, ONE kzgb200 context over all N GPUs
                    (kzgb200_opts.n_devices): fixed and per-GPU-scaled batches from one pinned host buffer.
Multi-GPU (main line): one process per GPU (torchrun), blobs sharded by rank, no data-path collective;
torch.distributed is used only for the barrier and the max-over-ranks of the step time.
"""
import argparse, ctypes, hashlib, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

R_MOD = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
BLOB = 131072

# Work model of SURVEY.md section 8(d): IMAD per Fp mul = 588, Fr mul = 264
IMAD_FP_MUL = 588
IMAD_FP_SQR = 456           # executed: the dedicated squaring is 222 wide products + 12 lo
CANONICAL_W = {"cells_proofs": 2605e6, "commit": 549e6, "blob_proof": 560e6, "verify_blob_batch": 8.2e6, "recover": 2701e6,
               "verify_cells": 1.79e6, "verify_cells_one_batch": 1.27e6}   # IMAD per unit, SURVEY 8(d)

METRIC = {"cells_proofs": "ComputeCellsAndKZGProofs blobs/s", "commit": "BlobToKZGCommitment blobs/s",
          "blob_proof": "ComputeBlobKZGProof blobs/s", "verify_blob_batch": "VerifyBlobKZGProofBatch blobs/s",
          "recover": "RecoverCellsAndComputeKZGProofs blobs/s", "verify_cells": "VerifyCellKZGProofBatch cells/s",
          "verify_cells_one_batch": "VerifyCellKZGProofBatch cells/s", "latency": "single-call latency (n = 1)"}
WORKLOAD_NAME = {"cells_proofs": "EIP-7594 ComputeCellsAndKZGProofs (FK20, 128 cells x 64 Fr), 1024 random blobs per GPU",
                 "commit": "EIP-4844 BlobToKZGCommitment, 4096 random blobs per GPU",
                 "blob_proof": "EIP-4844 ComputeBlobKZGProof, 4096 random blobs per GPU",
                 "verify_blob_batch": "EIP-4844 VerifyBlobKZGProofBatch, one RLC verdict over 4096 blobs per GPU",
                 "recover": "EIP-7594 RecoverCellsAndComputeKZGProofs, 64 of 128 cells (random pattern), 1024 blobs per GPU",
                 "verify_cells": "EIP-7594 VerifyCellKZGProofBatch, independent 128-cell batches, 128 x B cells per GPU",
                 "verify_cells_one_batch": "EIP-7594 VerifyCellKZGProofBatch, ONE verdict over all 128 x B cells per GPU (SURVEY 8d secondary shape)",
                 "latency": "one blob per call, every Context method"}
DEFAULT_B = {"cells_proofs": 1024, "commit": 4096, "blob_proof": 4096, "verify_blob_batch": 4096, "recover": 1024, "verify_cells": 4096,
             "verify_cells_one_batch": 4096, "latency": 1}
DTYPE = "u32 limbs (Fp 381-bit / Fr 255-bit Montgomery, integer)"


def unit_of(wl):
    return "cells/s" if wl.startswith("verify_cells") else "blobs/s"


def make_config(wl, B, world):
    """identical for both arms (--impl ours / reference): names the workload only"""
    per_unit = {"commit": BLOB, "blob_proof": BLOB + 48, "verify_blob_batch": BLOB + 96, "cells_proofs": BLOB, "recover": 64 * 2048,
                "verify_cells": 128 * 2144, "verify_cells_one_batch": 128 * 2144, "latency": BLOB}[wl]
    return {"workload": WORKLOAD_NAME[wl], "blobs_per_gpu_per_step": B,
            "l2_policy": "inputs larger than L2 (%.0f MB of input per step; the GPU arm also gathers from a multi-GB table at random)" % (B * per_unit / 1e6),
            "parallelism": "blob-sharded, %d process(es), no collective" % world}


def rand_blob(seed):
    out = bytearray(BLOB)
    sha = hashlib.sha256
    for j in range(4096):
        d = sha((seed + 32 * j).to_bytes(8, "big", signed=True)).digest()
        out[32 * j:32 * j + 32] = (int.from_bytes(d, "big") % R_MOD).to_bytes(32, "big")
    return bytes(out)


_blob_cache = {}


def make_blobs(first, count):
    """count distinct blobs; generated in parallel on the host cores (memoised per process: the extras reuse the same ranges)"""
    key = (first, count)
    if key not in _blob_cache:
        if len(_blob_cache) > 2:
            _blob_cache.clear()
        _blob_cache[key] = _make_blobs(first, count)
    return list(_blob_cache[key])


def _make_blobs(first, count):
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


# ======================================================================================================================
# CPU arm: the oracle port (reference's algorithmic structure, C++) on all host cores
# ======================================================================================================================
CPU_NOTE = "C++ restatement of the reference algorithm (oracle/), NOT gnark-crypto: no Go toolchain in this image"


class CpuSample:
    """A bounded sample of a workload for the oracle port: prepare() once (untimed), then run_pass() any number of times;
    one pass processes `units` units on `cores` host threads."""

    def __init__(self, workload, seconds_per_pass):
        import oracle_lib
        self.wl, self.o, self.L = workload, oracle_lib.get_oracle(), oracle_lib.lib()
        self.cores = os.cpu_count() or 1
        cores, o = self.cores, self.o
        if workload in ("commit", "cells_proofs"):
            self.kind = 0 if workload == "commit" else 1
            per = 0.012 if self.kind == 0 else 0.45                      # rough single-thread seconds per blob
            n = max(cores, int(seconds_per_pass * cores / per))
            n = min(n, 64 * cores)
            self.blobs = b"".join(make_blobs(1 << 30, min(n, 32)) * ((n + 31) // 32))[:n * BLOB]
            self.a = ctypes.create_string_buffer(n * (48 if self.kind == 0 else 262144))
            self.b = ctypes.create_string_buffer(n * 6144 if self.kind else 1)
            self.units, self.unit = n, "blobs/s"
            self.what = f"{n} blobs of the same generator per pass, blob-parallel"
            self.run_pass = self._pass_blobs
            return
        from concurrent.futures import ThreadPoolExecutor
        nb = 2 * cores
        blobs = make_blobs(1 << 30, nb)
        self.pool = ThreadPoolExecutor(cores)
        pool = self.pool
        cms = list(pool.map(lambda b: o.blob_to_kzg_commitment(b)[1], blobs))
        self.unit, per_call_units = "blobs/s", 1
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
                self.unit, per_call_units = "cells/s", 128
                what = "VerifyCellKZGProofBatch, one 128-cell verdict per call"
        self.n_calls = cores if workload == "verify_blob_batch" else nb
        self.call, self.units = call, self.n_calls * per_call_units
        self.what = f"{what}; {self.n_calls} calls per pass"
        self.run_pass = self._pass_calls

    def _pass_blobs(self):
        n = self.units
        rc = self.L.ko_parallel_blobs(self.o.ctx, self.kind, self.blobs, ctypes.c_size_t(n), self.cores, self.a, self.b)
        assert rc == 0

    def _pass_calls(self):
        assert all(st == 0 for st in self.pool.map(self.call, range(self.n_calls)))


CPU_CACHE = "/tmp/kzgb200_cpu_baseline_%s.json"


def cpu_baseline(workload, seconds_budget=10.0):
    """cpu_baseline object of the GPU arm's line.  If the reference arm ran on this box within the last hour (the driver runs it
    first), its number is reused instead of timing the same port twice (two numbers for one thing)."""
    try:
        c = json.load(open(CPU_CACHE % workload))
        if time.time() - c["when"] < 3600 and c["cores"] == (os.cpu_count() or 1):
            c["object"]["sample"] += " [reused from the --impl reference run on this box %.0f s earlier]" % (time.time() - c["when"])
            return c["object"]
    except Exception:
        pass
    s = CpuSample(workload, seconds_budget / 3)
    s.run_pass()                                                              # warm-up + correctness of the sample
    t = time.perf_counter(); passes = 0
    while passes < 2 or time.perf_counter() - t < seconds_budget / 2:
        s.run_pass(); passes += 1
    dt = time.perf_counter() - t
    return {"value": s.units * passes / dt, "unit": s.unit, "cores": s.cores, "kind": "port",
            "sample": f"{s.what}, {s.cores} host threads, {passes} passes in {dt:.1f} s; " + CPU_NOTE}


def run_reference(args):
    """--impl reference: W warm-up passes and EXACTLY K timed passes of a bounded sample (sized so the run ends within a few minutes)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    B = args.blobs or DEFAULT_B[wl]
    per_pass = min(4.0, max(0.5, 120.0 / (args.steps + args.warmup)))
    s = CpuSample(wl, per_pass)
    for _ in range(args.warmup):
        s.run_pass()
    t = time.perf_counter()
    for _ in range(args.steps):
        s.run_pass()
    dt = time.perf_counter() - t
    value = s.units * args.steps / dt
    base = {"value": value, "unit": s.unit, "cores": s.cores, "kind": "port",
            "sample": f"{s.what}, {s.cores} host threads, {args.steps} timed passes in {dt:.1f} s; " + CPU_NOTE}
    try:
        json.dump({"when": time.time(), "cores": s.cores, "object": dict(base)}, open(CPU_CACHE % wl, "w"))
    except Exception:
        pass
    line = {"impl": "reference", "metric": METRIC[wl], "value": value, "unit": s.unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 limbs (Fp 381-bit / Fr 255-bit Montgomery)", "data": "synthetic",
            "config": make_config(wl, B, int(os.environ.get("WORLD_SIZE", "1"))), "sample_units_per_step": s.units,
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": s.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ======================================================================================================================
# GPU arm
# ======================================================================================================================
class Work:
    """one workload on one context: step(dev) runs one pass on device (dev=True) or pinned host buffers"""
    pass


def make_work(ctx, wl, B, first_blob, torch, np, device, interleaved=False):
    """interleaved: host buffers come from kzgb200_host_alloc_interleaved (pages spread over the NUMA nodes) instead of torch's pinned pool"""
    import kzgb200
    L = ctx.L
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    SZ = ctypes.c_size_t
    dev = "cuda:%d" % device
    w = Work()
    w.host_ptrs = []

    def pinned(n, dt=torch.uint8):
        if not interleaved:
            return torch.empty(n, dtype=dt).pin_memory()
        nbytes = max(1, n) * torch.empty(0, dtype=dt).element_size()
        ptr = L.kzgb200_host_alloc_interleaved(nbytes)
        assert ptr, "kzgb200_host_alloc_interleaved failed"
        w.host_ptrs.append(ptr)
        return torch.frombuffer((ctypes.c_uint8 * nbytes).from_address(ptr), dtype=torch.uint8).view(dt)[:n]

    def pin(t):
        if not interleaved:
            return t.pin_memory()
        out = pinned(t.numel(), t.dtype)
        out.copy_(t.reshape(-1))
        return out

    def release():
        for ptr in w.host_ptrs:
            L.kzgb200_host_free(ctypes.c_void_p(ptr))
        w.host_ptrs.clear()
    w.release = release
    w.wl, w.B, w.units = wl, B, B
    blobs = make_blobs(first_blob, B)
    w.first_blob_bytes, w.last_blob_bytes = blobs[0], blobs[-1]
    h_blobs = pin(torch.frombuffer(bytearray(b"".join(blobs)), dtype=torch.uint8))
    del blobs
    d_blobs = h_blobs.to(dev)
    d_st = torch.zeros(max(B, 1), dtype=torch.int32, device=dev); h_st = pinned(max(B, 1), torch.int32)
    w.d_st = d_st
    w.keep = [h_blobs, d_blobs, d_st, h_st]
    import oracle_lib
    orc = oracle_lib.get_oracle

    def cut(t, i, size):
        return bytes(t[i * size:(i + 1) * size].cpu().numpy())

    if wl == "commit":
        d_o = torch.empty(48 * B, dtype=torch.uint8, device=dev); h_o = pinned(48 * B)
        def step(on_dev):
            i, o, s = (d_blobs, d_o, d_st) if on_dev else (h_blobs, h_o, h_st)
            ctx._check(L.kzgb200_blob_to_kzg_commitment(ctx.ctx, P(i), SZ(B), P(o), P(s)))
        w.h2d, w.d2h = B * BLOB, (48 + 4) * B
        w.self_check = lambda: bytes(h_o.numpy()) == bytes(d_o.cpu().numpy())
        w.oracle_check = lambda: all(cut(h_o, i, 48) == orc().blob_to_kzg_commitment(b)[1] for i, b in ((0, w.first_blob_bytes), (B - 1, w.last_blob_bytes)))
        w.host_check = lambda: w.oracle_check() and int(h_st.abs().sum().item()) == 0
    elif wl == "blob_proof":
        d_c = torch.empty(48 * B, dtype=torch.uint8, device=dev)
        ctx._check(L.kzgb200_blob_to_kzg_commitment(ctx.ctx, P(d_blobs), SZ(B), P(d_c), P(d_st)))
        h_c = pin(d_c.cpu())
        d_o = torch.empty(48 * B, dtype=torch.uint8, device=dev); h_o = pinned(48 * B)
        def step(on_dev):
            i, c_, o, s = (d_blobs, d_c, d_o, d_st) if on_dev else (h_blobs, h_c, h_o, h_st)
            ctx._check(L.kzgb200_compute_blob_kzg_proof(ctx.ctx, P(i), P(c_), SZ(B), P(o), P(s)))
        w.h2d, w.d2h = B * (BLOB + 48), (48 + 4) * B
        w.self_check = lambda: bytes(h_o.numpy()) == bytes(d_o.cpu().numpy())
        w.oracle_check = lambda: all(cut(h_o, i, 48) == orc().compute_blob_kzg_proof(b, cut(h_c, i, 48))[1] for i, b in ((0, w.first_blob_bytes), (B - 1, w.last_blob_bytes)))
    elif wl == "verify_blob_batch":
        d_c = torch.empty(48 * B, dtype=torch.uint8, device=dev); d_p = torch.empty(48 * B, dtype=torch.uint8, device=dev)
        ctx._check(L.kzgb200_blob_to_kzg_commitment(ctx.ctx, P(d_blobs), SZ(B), P(d_c), P(d_st)))
        ctx._check(L.kzgb200_compute_blob_kzg_proof(ctx.ctx, P(d_blobs), P(d_c), SZ(B), P(d_p), P(d_st)))
        h_c = pin(d_c.cpu()); h_p = pin(d_p.cpu())
        res = ctypes.c_int32(-1)
        def step(on_dev):
            i, c_, p_ = (d_blobs, d_c, d_p) if on_dev else (h_blobs, h_c, h_p)
            ctx._check(L.kzgb200_verify_blob_kzg_proof_batch(ctx.ctx, P(i), P(c_), P(p_), SZ(B), ctypes.byref(res)))
            assert res.value == 0, "valid batch rejected"
        def self_check():      # a corrupted proof must be rejected (Cfg2)
            bad = d_p.clone(); bad[17 * 48:18 * 48] = torch.frombuffer(bytearray(kzgb200.load_trusted_setup()[0][:48]), dtype=torch.uint8).to(dev)
            torch.cuda.synchronize()       # the library's streams do not wait for torch's (kzgb200.h: device inputs must be complete)
            ctx._check(L.kzgb200_verify_blob_kzg_proof_batch(ctx.ctx, P(d_blobs), P(d_c), P(bad), SZ(B), ctypes.byref(res)))
            return res.value == 1
        w.h2d, w.d2h = B * (BLOB + 96), 4
        w.self_check = self_check
        w.oracle_check = lambda: orc().verify_blob_kzg_proof(w.first_blob_bytes, cut(h_c, 0, 48), cut(h_p, 0, 48)) == 0 and \
            orc().verify_blob_kzg_proof(w.last_blob_bytes, cut(h_c, B - 1, 48), cut(h_p, 0, 48)) == 1
        w.keep += [d_c, d_p, h_c, h_p]
    else:
        # 7594 workloads need cells + proofs of the blobs
        d_cells = torch.empty(262144 * B, dtype=torch.uint8, device=dev); d_pr = torch.empty(6144 * B, dtype=torch.uint8, device=dev)
        h_cells = pinned(262144 * B); h_pr = pinned(6144 * B)
        def direct_ok(host=False):
            cells_t, pr_t = (h_cells, h_pr) if host else (d_cells, d_pr)
            ok = True
            for i, b in ((0, w.first_blob_bytes), (B - 1, w.last_blob_bytes)):
                est, ecells, eproofs = orc().compute_cells_and_kzg_proofs(b)
                ok = ok and est == 0 and cut(cells_t, i, 262144) == ecells and cut(pr_t, i, 6144) == eproofs
            return ok
        if wl == "cells_proofs":
            def step(on_dev):
                i, a, b, s = (d_blobs, d_cells, d_pr, d_st) if on_dev else (h_blobs, h_cells, h_pr, h_st)
                ctx._check(L.kzgb200_compute_cells_and_kzg_proofs(ctx.ctx, P(i), SZ(B), P(a), P(b), P(s)))
            w.h2d, w.d2h = B * BLOB, (262144 + 6144 + 4) * B
            w.self_check = lambda: bytes(h_pr.numpy()) == bytes(d_pr.cpu().numpy()) and bool(torch.equal(h_cells, d_cells.cpu()))
            w.oracle_check = direct_ok
            w.host_check = lambda: direct_ok(host=True) and int(h_st.abs().sum().item()) == 0     # outputs of the HOST-pointer path against the oracle
        else:
            # the inputs of recovery / cell verification are produced by THIS context (same tables as the timed calls)
            ctx._check(L.kzgb200_compute_cells_and_kzg_proofs(ctx.ctx, P(d_blobs), SZ(B), P(d_cells), P(d_pr), P(d_st)))
            assert int(d_st.abs().sum().item()) == 0
            if wl == "recover":
                ids = np.concatenate([np.sort(np.random.default_rng(first_blob + b).choice(128, 64, replace=False)) for b in range(B)]).astype(np.uint64)
                counts = np.full(B, 64, dtype=np.uint64)
                cells_np = d_cells.cpu().numpy().reshape(B, 128, 2048)
                sel = np.stack([cells_np[b, ids[64 * b:64 * b + 64].astype(np.int64)] for b in range(B)])      # [B,64,2048]
                h_in = pin(torch.from_numpy(np.ascontiguousarray(sel.reshape(-1)))); d_in = h_in.to(dev)
                d_oc = torch.empty(262144 * B, dtype=torch.uint8, device=dev); d_op = torch.empty(6144 * B, dtype=torch.uint8, device=dev)
                h_oc = pinned(262144 * B); h_op = pinned(6144 * B)
                idp = ids.ctypes.data_as(ctypes.c_void_p); cnp = counts.ctypes.data_as(ctypes.c_void_p)
                def step(on_dev):
                    i, a, b, s = (d_in, d_oc, d_op, d_st) if on_dev else (h_in, h_oc, h_op, h_st)
                    ctx._check(L.kzgb200_recover_cells_and_kzg_proofs(ctx.ctx, idp, cnp, P(i), SZ(B), P(a), P(b), P(s)))
                w.h2d, w.d2h = B * 64 * 2048, (262144 + 6144 + 4) * B
                w.self_check = lambda: bool(torch.equal(d_oc, d_cells)) and bool(torch.equal(d_op, d_pr)) and bool(torch.equal(h_oc, d_cells.cpu()))   # recovery == direct computation
                w.oracle_check = direct_ok
                w.keep += [ids, counts, h_in, d_in, d_oc, d_op, h_oc, h_op]
            else:   # verify_cells: B independent 128-cell batches, or one verdict
                d_cm1 = torch.empty(48 * B, dtype=torch.uint8, device=dev)
                ctx._check(L.kzgb200_blob_to_kzg_commitment(ctx.ctx, P(d_blobs), SZ(B), P(d_cm1), P(d_st)))
                N = 128 * B
                d_cm = d_cm1.view(B, 1, 48).expand(B, 128, 48).contiguous().view(-1)
                h_cm = pin(d_cm.cpu()); h_cells.copy_(d_cells); h_pr.copy_(d_pr)
                one = wl == "verify_cells_one_batch"
                idx = np.tile(np.arange(128, dtype=np.uint64), B)
                offs = np.array([0, N], dtype=np.uint64) if one else (np.arange(B + 1, dtype=np.uint64) * 128)
                nv = 1 if one else B
                d_res = torch.empty(nv, dtype=torch.int32, device=dev); h_res = pinned(nv, torch.int32)
                ip = idx.ctypes.data_as(ctypes.c_void_p); op_ = offs.ctypes.data_as(ctypes.c_void_p)
                def step(on_dev):
                    cm, ce, pr, rs = (d_cm, d_cells, d_pr, d_res) if on_dev else (h_cm, h_cells, h_pr, h_res)
                    ctx._check(L.kzgb200_verify_cell_kzg_proof_batch(ctx.ctx, P(cm), ip, P(ce), P(pr), SZ(N), op_, SZ(nv), P(rs)))
                w.units = N
                w.h2d, w.d2h = N * (48 + 2048 + 48), 4 * nv
                def self_check():
                    ok = int(d_res.abs().sum().item()) == 0 and int(h_res.abs().sum().item()) == 0
                    badc = d_cells.clone(); badc[5 * 2048 + 40] ^= 1                                   # corrupt one cell of batch 0
                    torch.cuda.synchronize()       # the library's streams do not wait for torch's (kzgb200.h: device inputs must be complete)
                    ctx._check(L.kzgb200_verify_cell_kzg_proof_batch(ctx.ctx, P(d_cm), ip, P(badc), P(d_pr), SZ(N), op_, SZ(nv), P(d_res)))
                    r = d_res.cpu()
                    return ok and int(r[0]) == 1 and int(r[1:].abs().sum()) == 0
                w.self_check = self_check
                def oracle_check():      # the oracle accepts what this context proved, and rejects the corrupted cell
                    cm = cut(d_cm1, 0, 48)
                    cl = [cut(d_cells, i, 2048) for i in range(128)]; pl = [cut(d_pr, i, 48) for i in range(128)]
                    bad = bytearray(cl[5]); bad[40] ^= 1
                    return direct_ok() and orc().verify_cell_kzg_proof_batch([cm] * 128, list(range(128)), cl, pl) == 0 and \
                        orc().verify_cell_kzg_proof_batch([cm] * 128, list(range(128)), cl[:5] + [bytes(bad)] + cl[6:], pl) == 1
                w.oracle_check = oracle_check
                w.host_check = lambda: int(h_res.abs().sum().item()) == 0 and oracle_check()
                w.keep += [idx, offs, d_cm, h_cm, d_res, h_res]
        w.keep += [d_cells, d_pr, h_cells, h_pr]
    w.step = step
    torch.cuda.synchronize()       # everything torch queued for the inputs is complete before the library's own streams read them
    return w


def time_work(ctx, w, steps, warmup, barrier, max_over_ranks, sampler=None):
    """W untimed + K timed device-resident steps, then the same K on pinned host buffers.  Returns a dict."""
    for _ in range(warmup):
        w.step(True)
    if sampler:
        sampler.start()
    l0 = ctx.info()["kernel_launches"]
    kms, dev_ms = {}, 0.0
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        w.step(True)
        dev_ms += ctx.last_device_ms()
        for k, v in ctx.last_kernel_ms().items():
            kms[k] = kms.get(k, 0.0) + v
    barrier()
    wall = time.perf_counter() - t0
    launches = ctx.info()["kernel_launches"] - l0
    clocks = sampler.stop() if sampler else None
    wall = max_over_ranks(wall); dev_ms = max_over_ranks(dev_ms)
    assert int(w.d_st.abs().sum().item()) == 0, "an item failed"
    w.step(False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        w.step(False)
    barrier()
    e2e_wall = max_over_ranks(time.perf_counter() - t0)
    return {"wall": wall, "e2e_wall": e2e_wall, "dev_ms": dev_ms, "kms": kms, "launches": launches, "clocks": clocks, "steps": steps}


def exact_pass(ctx, w, steps, barrier, max_over_ranks, world):
    """verify_cells with the optimistic first pass (one combined check over all verdicts, kzgb200_verify.cu) switched OFF: every verdict by
    its own random linear combination and pairing check -- what a call costs when the combined check does not pass (then on top of it)"""
    assert ctx.L.kzgb200_dbg_set_tunable(b"optimistic", 0) == 0
    try:
        w.step(True)
        kms = {}
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            w.step(True)
            for k, v in ctx.last_kernel_ms().items():
                kms[k] = kms.get(k, 0.0) + v
        barrier()
        wall = max_over_ranks(time.perf_counter() - t0)
    finally:
        ctx.L.kzgb200_dbg_set_tunable(b"optimistic", 1)
    return {"value": w.units * world * steps / wall, "unit": unit_of(w.wl), "ms_per_step": wall / steps * 1e3,
            "kernel_ms_per_step": {k: v / steps for k, v in kms.items() if v},
            "note": "tunable optimistic = 0: per-verdict checks only (round-1 behaviour); an input whose combined check fails pays the combined pass AND this one"}


def fmaheavy_from_profiles(kernel):
    """sm__pipe_fmaheavy_cycles_active of the committed ncu --set full capture of `kernel` (profiles/r02_ncu_pipes.json), or None"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_pipes.json"))).get(kernel)
    except Exception:
        return None


def roofline_for(wl, info, per, B, units, world_value_per_gpu, peaks, dev_ms_per_step):
    """roofline object for the dominant kernel (class) of a workload; `per` = kernel-class ms per step"""
    # peak_imad = device-wide 32-bit IMAD results/s (measured live, mad.lo); a full 32x32->64 product (IMAD.WIDE) counts
    # as 2 IMAD and issues at half that rate on the fmaheavy pipe: 148 SM x 32 lanes/clk (profiles/r02_pipe_probe.md)
    peak = peaks["mad_lo"]
    roof = {"bound": "int32-imad", "peak": peak / 1e12, "unit": "TIMAD/s",
            "peak_source": "measured live (kzgb200_bench_imad): mad.lo %.2f T lane-ops/s = the IMAD peak of SURVEY 8(d); an IMAD.WIDE counts as 2 IMAD and "
                           "issues at 32 lanes/clk/SM, so this peak is also 2 x the wide-multiply issue limit (profiles/r02_pipe_probe.md). Isolated chains "
                           "of the library's own field ops, measured live: Fp mul %.1f G/s, Fp sqr %.1f G/s"
                           % (peaks["mad_lo"] / 1e12, peaks["fp_mul"] / 1e9, peaks["fp_sqr"] / 1e9),
            "traffic": None,
            "whole_step_canonical": {"W_imad_per_unit": CANONICAL_W[wl], "achieved": world_value_per_gpu * CANONICAL_W[wl] / 1e12,
                                     "frac": world_value_per_gpu * CANONICAL_W[wl] / peak}}
    uses_commit = wl in ("commit", "blob_proof")
    uses_fk20 = wl in ("cells_proofs", "recover")
    # gathered mixed addition (g1.cuh g1_add_affine with the *Lazy policy): 6 mul + 2 sqr + ONE mul_add_mul (Y3 = R(Q-X3) - Y1*PPP under
    # one Montgomery reduction: 3 x 144 wide products + 12 = 876 IMAD instead of 2 x 588)
    IMAD_FP_MAM = 2 * 3 * 144 + 12
    IMAD_MIXED_ADD = 6 * IMAD_FP_MUL + 2 * IMAD_FP_SQR + IMAD_FP_MAM
    MULEQ_MIXED_ADD = 6.0 + 2.0 + 1.5                                         # in isolated-chain terms: a mul_add_mul ~ 1.5 products
    blended = MULEQ_MIXED_ADD / (7.5 / peaks["fp_mul"] + 2.0 / peaks["fp_sqr"])
    if uses_commit or uses_fk20:
        c_used, W_used = (info["commit_window"], info["commit_windows_per_scalar"]) if uses_commit else (info["fk20_window"], info["fk20_windows_per_scalar"])
        npts = 4096 if uses_commit else 8192
        msm_imad = npts * W_used * IMAD_MIXED_ADD
        msm_ms = per.get("msm", 0.0)
        achieved = msm_imad * B / (msm_ms * 1e-3) if msm_ms else None
        muls = npts * W_used * MULEQ_MIXED_ADD * B / (msm_ms * 1e-3) if msm_ms else None
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        roof.update({"kernel": "k_msm_fixed", "achieved": achieved / 1e12 if achieved else None, "frac": achieved / peak if achieved else None,
                     "frac_of_isolated_field_op_rate": muls / blended if muls else None,
                     "fmaheavy_pct_ncu": fmaheavy_from_profiles("k_msm_fixed"),
                     "work_model": "executed: %d pts x %d windows x (6 mul x 588 + 2 sqr x 456 + 1 two-product-one-reduction x 876) IMAD = %.0f M IMAD/blob in this kernel "
                                   "(round 1 executed 8 mul + 2 sqr = 5616 per addition; nominal 10 x 588: %.0f M); frac_of_isolated_field_op_rate = Fp mul-equivalents/s "
                                   "over the live-measured rate of a bare chain of the same mul : sqr mix"
                                   % (npts, W_used, msm_imad / 1e6, npts * W_used * 10 * IMAD_FP_MUL / 1e6),
                     "hbm_table_gather": {"achieved": npts * W_used * 96.0 * B / (msm_ms * 1e-3) / 1e9 if msm_ms else None, "unit": "GB/s", "peak": hbm_peak}})
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get("cells_proofs" if uses_fk20 else "commit")
            if tr and tr["window_bits"] == c_used:
                roof["traffic"] = tr["bytes_per_launch"]
                roof["traffic_note"] = "ncu dram bytes per 1024-blob launch of k_msm_fixed (%s); algorithmic gather bytes %d" % (tr.get("source", "profiles"), tr["algorithmic_gather_bytes_per_launch"])
        except Exception:
            pass
        return roof
    # verifier workloads: executed-work model of the two throughput kernels
    #   k_g1_check per point: sqrt by 4-bit windows (384 sqr + 97 mul) + on-curve test (3) + subgroup test
    #     (126 Jacobian doublings of 2 mul + 5 sqr, 10 additions of 11 mul + 3 sqr, 2 x cached Z^2, Z^3)
    #   k_vmsm_buckets per point: 96 windows x 15/16 non-zero digits x mixed addition (8 mul + 2 sqr)
    SQR = IMAD_FP_SQR
    DECODE = (384 * SQR + 100 * IMAD_FP_MUL) + 126 * (2 * IMAD_FP_MUL + 5 * SQR) + 10 * (11 * IMAD_FP_MUL + 3 * SQR) + 4 * IMAD_FP_MUL
    VMSM = 96 * 15 / 16 * IMAD_MIXED_ADD
    pts_decode = {"verify_cells": units + B, "verify_cells_one_batch": units + B, "verify_blob_batch": 2 * B}[wl]
    # large verdicts (and the optimistic combined pass of verify_cells): 32 of the 96 windows per point
    pts_vmsm = {"verify_cells": units * (32 / 96), "verify_cells_one_batch": units * (32 / 96), "verify_blob_batch": 2 * B * (64 / 96)}[wl]
    # the cell verifier's "decode" span (CUDA events on the main stream) also covers the side chain that runs beside k_g1_check on the aux
    # stream: the coset interpolation of the cells (448 Fr products of 264 IMAD per cell) -- counted, so that the fraction is of the whole span
    INTERP = 448 * 264
    side = units * INTERP if wl in ("verify_cells", "verify_cells_one_batch") and not per.get("fr") else 0
    models = {"decode": ("k_g1_check" + (" + k_cell_interp (side stream)" if side else ""), pts_decode * DECODE + side), "vmsm": ("k_vmsm_buckets (+ reduce, combine)", pts_vmsm * VMSM)}
    top = max((k for k in per if k in models and per[k]), key=lambda k: per[k], default=None)
    roof["kernel_classes"] = {k: {"kernel": models[k][0], "ms": per[k], "executed_imad": models[k][1],
                                  "achieved": models[k][1] / (per[k] * 1e-3) / 1e12, "frac": models[k][1] / (per[k] * 1e-3) / peak,
                                  "fmaheavy_pct_ncu": fmaheavy_from_profiles(models[k][0].split(" ")[0])}
                              for k in models if per.get(k)}
    if top:
        t = roof["kernel_classes"][top]
        roof.update({"kernel": t["kernel"], "achieved": t["achieved"], "frac": t["frac"], "fmaheavy_pct_ncu": t["fmaheavy_pct_ncu"],
                     "work_model": "executed IMAD of the dominant kernel class '%s' (%.1f of %.1f ms device time): %.0f k IMAD per decoded point, %.0f k per bucket-MSM point; "
                                   "latency-bound tails (hashing, single pairing) are listed in kernel_ms_per_step" % (top, per[top], dev_ms_per_step, DECODE / 1e3, VMSM / 1e3)})
    return roof


def latency_table(ctx, torch, np, reps=60, with_cpu=True):
    """p50 / p99 wall-clock of ONE-blob calls through HOST pointers (the reference's calling pattern: prove.go:13, api_eip7594.go:28, ...),
    the device-side time of the same call, and the oracle port's single-thread time of the same call."""
    import oracle_lib
    L = ctx.L
    SZ = ctypes.c_size_t
    o = oracle_lib.get_oracle()
    blob = rand_blob(123 << 20)
    st, cm = ctx.blob_to_kzg_commitment(blob); assert st == 0
    z = (0x1234567 ** 7 % R_MOD).to_bytes(32, "big")
    st, pz, y = ctx.compute_kzg_proof(blob, z); assert st == 0
    st, bp = ctx.compute_blob_kzg_proof(blob, cm); assert st == 0
    st, cells, proofs = ctx.compute_cells_and_kzg_proofs(blob); assert st == 0
    ids = list(range(0, 128, 2))
    half = b"".join(cells[2048 * i:2048 * i + 2048] for i in ids)
    pin = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8).pin_memory()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    h_blob, h_cm, h_z, h_y, h_pz, h_bp, h_cells, h_proofs, h_half = map(pin, (blob, cm, z, y, pz, bp, cells, proofs, half))
    h_cm128 = pin(cm * 128)
    o48 = torch.empty(48, dtype=torch.uint8).pin_memory(); o32 = torch.empty(32, dtype=torch.uint8).pin_memory()
    oc = torch.empty(262144, dtype=torch.uint8).pin_memory(); op = torch.empty(6144, dtype=torch.uint8).pin_memory()
    st1 = torch.zeros(1, dtype=torch.int32).pin_memory()
    res = ctypes.c_int32()
    idx = np.arange(128, dtype=np.uint64); offs = np.array([0, 128], dtype=np.uint64)
    ids_np = np.array(ids, dtype=np.uint64); cnt = np.array([64], dtype=np.uint64)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    cl = [cells[2048 * i:2048 * i + 2048] for i in range(128)]; pl = [proofs[48 * i:48 * i + 48] for i in range(128)]
    calls = {
        "BlobToKZGCommitment": (lambda: L.kzgb200_blob_to_kzg_commitment(ctx.ctx, P(h_blob), SZ(1), P(o48), P(st1)), lambda: o.blob_to_kzg_commitment(blob)),
        "ComputeKZGProof": (lambda: L.kzgb200_compute_kzg_proof(ctx.ctx, P(h_blob), P(h_z), SZ(1), P(o48), P(o32), P(st1)), lambda: o.compute_kzg_proof(blob, z)),
        "ComputeBlobKZGProof": (lambda: L.kzgb200_compute_blob_kzg_proof(ctx.ctx, P(h_blob), P(h_cm), SZ(1), P(o48), P(st1)), lambda: o.compute_blob_kzg_proof(blob, cm)),
        "VerifyKZGProof": (lambda: L.kzgb200_verify_kzg_proof(ctx.ctx, P(h_cm), P(h_z), P(h_y), P(h_pz), SZ(1), P(st1)), lambda: o.verify_kzg_proof(cm, z, y, pz)),
        "VerifyBlobKZGProof": (lambda: L.kzgb200_verify_blob_kzg_proof(ctx.ctx, P(h_blob), P(h_cm), P(h_bp), SZ(1), P(st1)), lambda: o.verify_blob_kzg_proof(blob, cm, bp)),
        "VerifyBlobKZGProofBatch(1)": (lambda: L.kzgb200_verify_blob_kzg_proof_batch(ctx.ctx, P(h_blob), P(h_cm), P(h_bp), SZ(1), ctypes.byref(res)),
                                       lambda: o.verify_blob_kzg_proof_batch([blob], [cm], [bp])),
        "ComputeCells": (lambda: L.kzgb200_compute_cells(ctx.ctx, P(h_blob), SZ(1), P(oc), P(st1)), lambda: o.compute_cells(blob)),
        "ComputeCellsAndKZGProofs": (lambda: L.kzgb200_compute_cells_and_kzg_proofs(ctx.ctx, P(h_blob), SZ(1), P(oc), P(op), P(st1)), lambda: o.compute_cells_and_kzg_proofs(blob)),
        "RecoverCellsAndComputeKZGProofs(64 cells)": (lambda: L.kzgb200_recover_cells_and_kzg_proofs(ctx.ctx, vp(ids_np), vp(cnt), P(h_half), SZ(1), P(oc), P(op), P(st1)),
                                                      lambda: o.recover_cells_and_kzg_proofs(ids, [cl[i] for i in ids])),
        "VerifyCellKZGProofBatch(128 cells)": (lambda: L.kzgb200_verify_cell_kzg_proof_batch(ctx.ctx, P(h_cm128), vp(idx), P(h_cells), P(h_proofs), SZ(128), vp(offs), SZ(1), P(st1)),
                                               lambda: o.verify_cell_kzg_proof_batch([cm] * 128, list(range(128)), cl, pl)),
    }
    out = {}
    for name, (gpu_call, cpu_call) in calls.items():
        for _ in range(5):
            assert gpu_call() == 0
        assert int(st1[0]) == 0 and (name != "VerifyBlobKZGProofBatch(1)" or res.value == 0), name
        ts, dms = [], []
        for _ in range(reps):
            t = time.perf_counter(); rc = gpu_call(); ts.append(time.perf_counter() - t); dms.append(ctx.last_device_ms())
            assert rc == 0
        ts.sort(); dms.sort()
        row = {"p50_ms": ts[len(ts) // 2] * 1e3, "p99_ms": ts[min(len(ts) - 1, int(len(ts) * 0.99))] * 1e3, "device_ms_p50": dms[len(dms) // 2]}
        if with_cpu:
            t = time.perf_counter(); r = cpu_call(); row["oracle_port_1thread_ms"] = (time.perf_counter() - t) * 1e3
            assert (r if isinstance(r, int) else r[0]) == 0, name
        out[name] = row
    if "ComputeCellsAndKZGProofs" in out:      # what the call computed is what the oracle computes
        assert bytes(op.numpy()) == proofs and bytes(oc.numpy()) == cells == o.compute_cells_and_kzg_proofs(blob)[1]
    return out


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
    ap.add_argument("--coresident", default="13,13", help="commit,fk20 windows of the co-resident context of `extras`")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.steps <= 0:
        args.steps = {"verify_blob_batch": 25, "verify_cells": 10, "verify_cells_one_batch": 10}.get(args.workload, 5)
    if args.impl == "reference":
        if args.workload == "latency":
            print(json.dumps({"impl": "reference", "unavailable": "the latency table already carries the oracle port's single-thread time per method"}))
            return
        return run_reference(args)

    # stdout carries exactly ONE JSON line: everything libraries print to fd 1 (e.g. "NCCL version ..." with NCCL_DEBUG
    # set on the box) is sent to stderr, and the result line is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import kzgb200, sharding
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
    # a CPU-side barrier for the section where rank 0 alone drives ALL GPUs from one process: a rank waiting in an NCCL barrier spins in a
    # kernel on its GPU, which would time-slice against rank 0's work on that GPU
    cpu_group = dist.new_group(backend="gloo") if dist else None

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

    def emit(line):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    dbg = kzgb200.Debug()
    if wl == "latency":
        ctx = kzgb200.Context(device=local, commit_window=13, fk20_window=13)
        tab = latency_table(ctx, torch, np, reps=200, with_cpu=not args.no_cpu_baseline)
        info = ctx.info(); ctx.close()
        if rank == 0:
            emit({"metric": METRIC[wl], "value": tab["ComputeCellsAndKZGProofs"]["p50_ms"], "unit": "ms (p50, ComputeCellsAndKZGProofs, one blob, host pointers)",
                  "n_gpus": 1, "steps": 200, "warmup": 5, "ms_per_step": tab["ComputeCellsAndKZGProofs"]["p50_ms"], "higher_is_better": False, "scaling": "weak",
                  "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": make_config(wl, 1, world),
                  "engine": {"commit_window": info["commit_window"], "fk20_window": info["fk20_window"]}, "latency_n1": tab})
        return

    # tables: only the one this workload needs gets the big window
    uses_commit = wl in ("commit", "blob_proof")
    uses_fk20 = wl in ("cells_proofs", "recover")
    cw = args.commit_window if uses_commit else 8
    fw = args.fk20_window if uses_fk20 else 8
    ctx = kzgb200.Context(device=local, commit_window=cw, fk20_window=fw)
    info0 = ctx.info()
    w = make_work(ctx, wl, B, rank * B, torch, np, local)
    sampler = ClockSampler(local)
    r = time_work(ctx, w, args.steps, args.warmup, barrier, max_over_ranks, sampler)
    assert w.self_check(), "self-check failed (host path vs device path / accept-reject)"
    exact = exact_pass(ctx, w, args.steps, barrier, max_over_ranks, world) if wl == "verify_cells" else None
    oracle_ok = bool(w.oracle_check())
    assert oracle_ok, "outputs differ from the CPU oracle"
    unit = unit_of(wl)
    total_units = w.units * world * args.steps
    value = total_units / r["wall"]
    peaks = {m: dbg.imad_peak(local, i) for i, m in ((0, "mad_lo"), (3, "fp_mul"), (4, "fp_sqr"))}
    per = {k: v / args.steps for k, v in r["kms"].items() if v}
    roof = roofline_for(wl, info0, per, B, w.units, value / world, peaks, r["dev_ms"] / args.steps)
    cfg = make_config(wl, B, world)
    engine = {"commit_window": info0["commit_window"], "fk20_window": info0["fk20_window"], "commit_table_bytes": info0["commit_table_bytes"],
              "fk20_table_bytes": info0["fk20_table_bytes"], "table_used": "commit" if uses_commit else "fk20" if uses_fk20 else "none (verifier)",
              "note": "only the table this workload gathers from is built at the quoted window; extras.coresident is ONE context serving all three metrics"}
    line = {
        "metric": METRIC[wl], "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["wall"] / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic", "config": cfg, "engine": engine,
        "device_ms_per_step": r["dev_ms"] / args.steps, "kernel_ms_per_step": per,
        "gpu_launches": r["launches"], "clocks": r["clocks"],
        "e2e": {"value": total_units / r["e2e_wall"], "unit": unit, "h2d_bytes_per_step": w.h2d, "d2h_bytes_per_step": w.d2h},
        "roofline": roof, "oracle_check": "ok" if oracle_ok else "MISMATCH",
        "oracle_check_what": "first and last item of the batch recomputed by the CPU oracle after the timed region (and, for the verifiers, accept + corrupted-input reject on both)",
    }
    if exact:
        line["per_verdict_pass"] = exact
    del w
    ctx.close()
    torch.cuda.empty_cache()

    extras = {}
    if wl == "cells_proofs" and not args.no_extras:
        cwc, fwc = (int(x) for x in args.coresident.split(","))
        cctx = kzgb200.Context(device=local, commit_window=cwc, fk20_window=fwc)
        ci = cctx.info()
        co = {"context": {"commit_window": ci["commit_window"], "fk20_window": ci["fk20_window"], "commit_table_bytes": ci["commit_table_bytes"],
                          "fk20_table_bytes": ci["fk20_table_bytes"], "init_ms": ci["init_ms"]},
              "steps": 3, "warmup": 3, "note": "one context, both tables resident together; weak scaling (per-GPU batch) like the main line"}
        strong = {}
        for sub, subB in (("cells_proofs", 1024), ("commit", 4096), ("verify_cells", 4096)):
            sw = make_work(cctx, sub, subB, rank * subB, torch, np, local)
            sr = time_work(cctx, sw, 3, 3, barrier, max_over_ranks)
            assert sw.self_check() and sw.oracle_check(), "co-resident context: " + sub
            sper = {k: v / 3 for k, v in sr["kms"].items() if v}
            sval = sw.units * world * 3 / sr["wall"]
            co[sub] = {"metric": METRIC[sub], "value": sval, "unit": unit_of(sub), "e2e": sw.units * world * 3 / sr["e2e_wall"], "ms_per_step": sr["wall"] / 3 * 1e3,
                       "kernel_ms_per_step": sper, "oracle_check": "ok", **({"per_verdict_pass": exact_pass(cctx, sw, 3, barrier, max_over_ranks, world)} if sub == "verify_cells" else {}),
                       "roofline": {k: v for k, v in roofline_for(sub, ci, sper, subB, sw.units, sval / world, peaks, sr["dev_ms"] / 3).items()
                                    if k in ("kernel", "achieved", "peak", "unit", "frac", "frac_of_isolated_field_op_rate", "fmaheavy_pct_ncu", "kernel_classes")}}
            del sw
            torch.cuda.empty_cache()
            if world > 1:
                # STRONG scaling: the same TOTAL batch as N = 1, cut over the ranks (BASELINE.json configs[2], [4]: "sharded 1/2/4/8")
                lo, hi = sharding.shard_range(subB, world, rank)
                sw = make_work(cctx, sub, hi - lo, lo, torch, np, local)
                sr = time_work(cctx, sw, 5, 3, barrier, max_over_ranks)
                tot = (subB * 128 if sub == "verify_cells" else subB) * 5
                strong[sub] = {"total_units_per_step": tot // 5, "per_rank": hi - lo, "value": tot / sr["wall"], "e2e": tot / sr["e2e_wall"], "unit": unit_of(sub),
                               "ms_per_step": sr["wall"] / 5 * 1e3, "steps": 5}
                del sw
                torch.cuda.empty_cache()
        extras["coresident"] = co
        if strong:
            extras["strong_scaling"] = dict(strong, note="fixed total batch over %d ranks (one process per GPU), max-over-ranks time" % world)
        if rank == 0:
            extras["latency_n1"] = dict(latency_table(cctx, torch, np, reps=60, with_cpu=not args.no_cpu_baseline and world == 1),
                                        note="one blob per call through pinned host pointers on the co-resident context; wall-clock p50 / p99 of 60 calls, device_ms = CUDA events "
                                             "around the kernels; oracle_port_1thread_ms = the CPU port's single call (N = 1 runs only)")
        barrier()
        cctx.close()
        torch.cuda.empty_cache()
        if world > 1:
            # ONE process, ONE context over all GPUs (kzgb200_opts.n_devices): rank 0 alone, the other ranks have released their tables and wait
            # on the CPU (gloo), so that nothing of theirs runs on the GPUs meanwhile
            barrier()
            if rank == 0:
                try:
                    mctx = kzgb200.Context(commit_window=cwc, fk20_window=fwc, devices=list(range(world)), lanes=2)
                    ip = {"context": {"devices": world, "commit_window": cwc, "fk20_window": fwc, "init_ms": mctx.info()["init_ms"]}}
                    devs = (ctypes.c_int * world)(*range(world))
                    for inter in (0, 1):
                        g = ctypes.c_double()
                        if mctx.L.kzgb200_dbg_h2d_bandwidth(devs, world, ctypes.c_size_t(256 << 20), inter, ctypes.byref(g)) == 0:
                            ip["h2d_ceiling_GBps_%s" % ("interleaved" if inter else "plain_pinned")] = g.value
                    for name, sub, subB in (("cells_proofs_fixed_1024", "cells_proofs", 1024), ("cells_proofs_1024_per_gpu", "cells_proofs", 1024 * world),
                                            ("verify_cells_fixed_524288", "verify_cells", 4096), ("commit_fixed_4096", "commit", 4096)):
                        sw = make_work(mctx, sub, subB, 0, torch, np, 0, interleaved=True)
                        sw.step(False); sw.step(False)
                        t0 = time.perf_counter()
                        for _ in range(4):
                            sw.step(False)
                        dt = (time.perf_counter() - t0) / 4
                        assert sw.host_check(), "in-process multi-GPU: " + name         # host outputs of the sharded call against the CPU oracle
                        ip[name] = {"e2e": sw.units / dt, "unit": unit_of(sub), "ms_per_step": dt * 1e3, "h2d_GBps": sw.h2d / dt / 1e9, "d2h_GBps": sw.d2h / dt / 1e9,
                                    "steps": 4, "oracle_check": "ok"}
                        sw.keep.clear(); sw.release()
                        del sw
                        torch.cuda.empty_cache()
                    ip["note"] = "host buffers only (ONE NUMA-interleaved pinned buffer per array, sharded inside the library over the GPUs); e2e = units / wall-clock of the C-ABI call; " \
                                 "h2d_ceiling = all GPUs copying 256 MB each from one host buffer at the same time"
                    mctx.close()
                    extras["in_process"] = ip
                except Exception as e:      # noqa: BLE001  (never lose the main line to an extra)
                    extras["in_process"] = {"error": repr(e)}
            torch.cuda.synchronize()
            dist.barrier(group=cpu_group)
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    if extras:
        line["extras"] = extras
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(wl)
    emit(line)


if __name__ == "__main__":
    main()
