"""ctypes binding of libkzgb200.so + a host-side mirror of the reference's `Context`.

The reference's boundary is the Go method set of *goethkzg.Context (api.go:17-28, prove.go,
verify.go, api_eip7594.go, api_eip.go).  No Go toolchain exists in this image, so this module is
the executable stand-in for the cgo shim shown in INTEGRATION.md: same method names (snake_case),
same argument meaning, and the same three-way outcome convention the reference's tests use
(nil / kzg.ErrVerifyOpeningProof / any other error  ==  OK / VERIFY_FAILED / other status).

There is NO CPU fallback: if libkzgb200.so is missing or no CUDA device is usable, construction
raises.
"""
import ctypes, os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libkzgb200.so")
SETUP = os.path.join(HERE, "data", "trusted_setup_4096.bin")

BYTES_PER_BLOB = 131072
BYTES_PER_CELL = 2048
CELLS_PER_EXT_BLOB = 128

OK, VERIFY_FAILED, NON_CANONICAL_SCALAR, BAD_G1_ENCODING, NOT_ON_CURVE, NOT_IN_SUBGROUP = 0, 1, 2, 3, 4, 5
LENGTH_MISMATCH, BAD_CELL_INDEX, CELL_IDS_NOT_ASCENDING, NOT_ENOUGH_CELLS, BAD_ROW_INDEX = 6, 7, 8, 9, 10
ERR_ARGS, ERR_SETUP, ERR_CUDA = 11, 12, 100


class KzgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"kzgb200 error {code}: {msg}")
        self.code = code


class Opts(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int), ("commit_window", ctypes.c_int), ("fk20_window", ctypes.c_int),
                ("n_devices", ctypes.c_int), ("devices", ctypes.POINTER(ctypes.c_int)), ("lanes", ctypes.c_int),
                ("reserved", ctypes.c_int * 1)]


class Info(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int), ("sm_count", ctypes.c_int),
                ("commit_window", ctypes.c_int), ("commit_windows_per_scalar", ctypes.c_int),
                ("fk20_window", ctypes.c_int), ("fk20_windows_per_scalar", ctypes.c_int),
                ("commit_table_bytes", ctypes.c_uint64), ("fk20_table_bytes", ctypes.c_uint64),
                ("init_ms", ctypes.c_double), ("kernel_launches", ctypes.c_uint64),
                ("n_devices", ctypes.c_int), ("lanes_per_device", ctypes.c_int)]


_lib = None

EXPORTS = [
    "kzgb200_ctx_new", "kzgb200_ctx_free", "kzgb200_last_error", "kzgb200_host_alloc", "kzgb200_host_alloc_interleaved", "kzgb200_host_free",
    "kzgb200_blob_to_kzg_commitment", "kzgb200_get_info", "kzgb200_last_device_ms",
    "kzgb200_compute_cells", "kzgb200_compute_cells_and_kzg_proofs", "kzgb200_last_kernel_ms",
    "kzgb200_compute_kzg_proof", "kzgb200_compute_blob_kzg_proof", "kzgb200_recover_cells_and_kzg_proofs",
    "kzgb200_verify_kzg_proof", "kzgb200_verify_blob_kzg_proof", "kzgb200_verify_blob_kzg_proof_batch",
    "kzgb200_verify_cell_kzg_proof_batch",
    "kzgb200_parse_trusted_setup_json", "kzgb200_ctx_new_from_json", "kzgb200_check_trusted_setup",
    "kzgb200_check_g1_points", "kzgb200_check_scalars",
]


def load_library():
    """Loads the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise KzgError(ERR_CUDA, f"{SO} not built: run `python go-eth-kzg_b200/build.py`")
        L = ctypes.CDLL(SO)
        L.kzgb200_last_error.restype = ctypes.c_char_p
        L.kzgb200_host_alloc.restype = ctypes.c_void_p
        L.kzgb200_host_alloc.argtypes = [ctypes.c_size_t]
        L.kzgb200_host_alloc_interleaved.restype = ctypes.c_void_p
        L.kzgb200_host_alloc_interleaved.argtypes = [ctypes.c_size_t]
        L.kzgb200_host_free.argtypes = [ctypes.c_void_p]
        L.kzgb200_last_device_ms.restype = ctypes.c_double
        L.kzgb200_last_device_ms.argtypes = [ctypes.c_void_p]
        L.kzgb200_ctx_free.argtypes = [ctypes.c_void_p]
        _lib = L
    return _lib


def load_trusted_setup(path=SETUP):
    raw = open(path, "rb").read()
    n = 4096 * 48
    return raw[:n], raw[n:2 * n], raw[2 * n:]


def parse_trusted_setup_json(text):
    """JSONTrustedSetup text (trusted_setup.go:23-27) -> (g1_monomial, g1_lagrange, g2_monomial) flat bytes; host only"""
    L = load_library()
    raw = text.encode() if isinstance(text, str) else bytes(text)
    m = ctypes.create_string_buffer(4096 * 48); l = ctypes.create_string_buffer(4096 * 48); g2 = ctypes.create_string_buffer(4096 * 96)
    n = ctypes.c_size_t()
    rc = L.kzgb200_parse_trusted_setup_json(raw, ctypes.c_size_t(len(raw)), m, l, g2, ctypes.c_size_t(4096), ctypes.byref(n))
    if rc != OK:
        raise KzgError(rc, L.kzgb200_last_error().decode())
    return m.raw, l.raw, g2.raw[:96 * n.value]


def check_trusted_setup(g1_lagrange, g1_monomial, g2_monomial, device=0):
    """CheckTrustedSetupIsWellFormed (trusted_setup.go:45-83) on the GPU -> (status, index of the first bad point)"""
    L = load_library()
    res = ctypes.c_int32(); bad = ctypes.c_size_t()
    rc = L.kzgb200_check_trusted_setup(device, g1_lagrange, ctypes.c_size_t(len(g1_lagrange) // 48), g1_monomial, ctypes.c_size_t(len(g1_monomial) // 48),
                                       g2_monomial, ctypes.c_size_t(len(g2_monomial) // 96), ctypes.byref(res), ctypes.byref(bad))
    if rc != OK:
        raise KzgError(rc, L.kzgb200_last_error().decode())
    return res.value, bad.value


def _ptr(x):
    """bytes / ctypes buffer / int address -> c_void_p"""
    if isinstance(x, int):
        return ctypes.c_void_p(x)
    if isinstance(x, (bytes, bytearray)):
        return ctypes.cast(ctypes.c_char_p(bytes(x)), ctypes.c_void_p)
    return ctypes.cast(x, ctypes.c_void_p)


class Context:
    """Mirror of goethkzg.Context.  NewContext4096Secure (api.go:53) == Context()."""

    def __init__(self, device=0, commit_window=0, fk20_window=0, setup_path=SETUP, setup_json=None, devices=None, lanes=0, setup=None):
        """setup_json: text of a JSONTrustedSetup (NewContext4096(&parsed), api.go:90); setup: (g1_monomial, g1_lagrange,
        g2_monomial) flat bytes; default: the packed mainnet setup.  devices: list of CUDA ordinals, or "all" (one replica of
        the tables per GPU, batched host-buffer calls are sharded over them); lanes: concurrent calls per GPU (0 = default)."""
        L = load_library()
        self.L = L
        opts = Opts(device=device, commit_window=commit_window, fk20_window=fk20_window, lanes=lanes)
        if devices == "all":
            opts.n_devices = -1
        elif devices:
            self._devs = (ctypes.c_int * len(devices))(*devices)
            opts.n_devices = len(devices)
            opts.devices = ctypes.cast(self._devs, ctypes.POINTER(ctypes.c_int))
        ctx = ctypes.c_void_p()
        if setup_json is not None:
            raw = setup_json.encode() if isinstance(setup_json, str) else bytes(setup_json)
            rc = L.kzgb200_ctx_new_from_json(raw, ctypes.c_size_t(len(raw)), ctypes.byref(opts), ctypes.byref(ctx))
        else:
            m, l, g2 = setup if setup is not None else load_trusted_setup(setup_path)
            rc = L.kzgb200_ctx_new(m, l, g2, ctypes.c_size_t(len(g2) // 96), ctypes.byref(opts), ctypes.byref(ctx))
        if rc != OK:
            raise KzgError(rc, L.kzgb200_last_error().decode())
        self.ctx = ctx

    def close(self):
        if getattr(self, "ctx", None):
            self.L.kzgb200_ctx_free(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise KzgError(rc, self.L.kzgb200_last_error().decode())

    def info(self):
        i = Info()
        self._check(self.L.kzgb200_get_info(self.ctx, ctypes.byref(i)))
        return {f[0]: getattr(i, f[0]) for f in Info._fields_}

    def last_device_ms(self):
        return self.L.kzgb200_last_device_ms(self.ctx)

    KERNEL_CLASSES = ["fr", "msm", "g1fft", "finalize", "verify", "decode", "vmsm", "pairing"]

    def last_kernel_ms(self):
        arr = (ctypes.c_double * 8)()
        self._check(self.L.kzgb200_last_kernel_ms(self.ctx, arr))
        return {k: arr[i] for i, k in enumerate(self.KERNEL_CLASSES)}

    # ---- raw batched calls on caller-provided addresses (host or device) -------------------
    def raw_blob_to_kzg_commitment(self, blobs_ptr, n, out_ptr, status_ptr):
        self._check(self.L.kzgb200_blob_to_kzg_commitment(self.ctx, _ptr(blobs_ptr), ctypes.c_size_t(n), _ptr(out_ptr), _ptr(status_ptr)))

    def raw_compute_cells_and_kzg_proofs(self, blobs_ptr, n, cells_ptr, proofs_ptr, status_ptr):
        self._check(self.L.kzgb200_compute_cells_and_kzg_proofs(self.ctx, _ptr(blobs_ptr), ctypes.c_size_t(n), _ptr(cells_ptr),
                                                                _ptr(proofs_ptr), _ptr(status_ptr)))

    def raw_compute_cells(self, blobs_ptr, n, cells_ptr, status_ptr):
        self._check(self.L.kzgb200_compute_cells(self.ctx, _ptr(blobs_ptr), ctypes.c_size_t(n), _ptr(cells_ptr), _ptr(status_ptr)))

    # ---- batched, bytes in / bytes out ------------------------------------------------------
    def compute_cells_and_kzg_proofs_batch(self, blobs, proofs=True):
        n = len(blobs)
        if any(len(b) != BYTES_PER_BLOB for b in blobs):
            raise KzgError(LENGTH_MISMATCH, "blob must be 131072 bytes")
        cells = ctypes.create_string_buffer(262144 * max(n, 1))
        st = (ctypes.c_int32 * max(n, 1))()
        if proofs:
            pr = ctypes.create_string_buffer(6144 * max(n, 1))
            self.raw_compute_cells_and_kzg_proofs(b"".join(blobs), n, cells, pr, st)
            craw, praw = cells.raw, pr.raw
            return [(st[i], craw[262144 * i:262144 * (i + 1)], praw[6144 * i:6144 * (i + 1)]) for i in range(n)]
        self.raw_compute_cells(b"".join(blobs), n, cells, st)
        craw = cells.raw
        return [(st[i], craw[262144 * i:262144 * (i + 1)]) for i in range(n)]

    def blob_to_kzg_commitment_batch(self, blobs):
        n = len(blobs)
        if any(len(b) != BYTES_PER_BLOB for b in blobs):
            raise KzgError(LENGTH_MISMATCH, "blob must be 131072 bytes")
        out = ctypes.create_string_buffer(48 * max(n, 1))
        st = (ctypes.c_int32 * max(n, 1))()
        self.raw_blob_to_kzg_commitment(b"".join(blobs), n, out, st)
        oraw = out.raw
        return [(st[i], oraw[48 * i:48 * i + 48]) for i in range(n)]

    def compute_kzg_proof_batch(self, blobs, zs):
        n = len(blobs)
        proofs = ctypes.create_string_buffer(48 * max(n, 1))
        ys = ctypes.create_string_buffer(32 * max(n, 1))
        st = (ctypes.c_int32 * max(n, 1))()
        self._check(self.L.kzgb200_compute_kzg_proof(self.ctx, _ptr(b"".join(blobs)), _ptr(b"".join(zs)), ctypes.c_size_t(n),
                                                     _ptr(proofs), _ptr(ys), _ptr(st)))
        praw, yraw = proofs.raw, ys.raw
        return [(st[i], praw[48 * i:48 * i + 48], yraw[32 * i:32 * i + 32]) for i in range(n)]

    def compute_blob_kzg_proof_batch(self, blobs, commitments):
        n = len(blobs)
        proofs = ctypes.create_string_buffer(48 * max(n, 1))
        st = (ctypes.c_int32 * max(n, 1))()
        self._check(self.L.kzgb200_compute_blob_kzg_proof(self.ctx, _ptr(b"".join(blobs)), _ptr(b"".join(commitments)), ctypes.c_size_t(n),
                                                          _ptr(proofs), _ptr(st)))
        praw = proofs.raw
        return [(st[i], praw[48 * i:48 * i + 48]) for i in range(n)]

    def recover_cells_and_kzg_proofs_batch(self, ids_list, cells_list, proofs=True):
        """ids_list[i] / cells_list[i]: the cell ids and 2048-byte cells provided for blob i"""
        n = len(ids_list)
        flat_ids = [x for ids in ids_list for x in ids]
        ids = (ctypes.c_uint64 * max(len(flat_ids), 1))(*flat_ids)
        counts = (ctypes.c_uint64 * max(n, 1))(*[len(x) for x in ids_list])
        cells = b"".join(c for cl in cells_list for c in cl)
        oc = ctypes.create_string_buffer(262144 * max(n, 1))
        op = ctypes.create_string_buffer(6144 * max(n, 1)) if proofs else None
        st = (ctypes.c_int32 * max(n, 1))()
        self._check(self.L.kzgb200_recover_cells_and_kzg_proofs(self.ctx, ids, counts, _ptr(cells) if cells else None, ctypes.c_size_t(n),
                                                                _ptr(oc), _ptr(op) if proofs else None, _ptr(st)))
        craw = oc.raw
        praw = op.raw if proofs else None
        return [(st[i], craw[262144 * i:262144 * (i + 1)]) + ((praw[6144 * i:6144 * (i + 1)],) if proofs else ()) for i in range(n)]

    def verify_kzg_proof_batch(self, commitments, zs, ys, proofs):
        """n independent VerifyKZGProof checks -> list of statuses"""
        n = len(commitments)
        st = (ctypes.c_int32 * max(n, 1))()
        self._check(self.L.kzgb200_verify_kzg_proof(self.ctx, _ptr(b"".join(commitments)), _ptr(b"".join(zs)), _ptr(b"".join(ys)),
                                                    _ptr(b"".join(proofs)), ctypes.c_size_t(n), st))
        return list(st[:n])

    def verify_blob_kzg_proof_each(self, blobs, commitments, proofs):
        """n independent VerifyBlobKZGProof checks (== VerifyBlobKZGProofBatchPar's work) -> statuses"""
        n = len(blobs)
        st = (ctypes.c_int32 * max(n, 1))()
        self._check(self.L.kzgb200_verify_blob_kzg_proof(self.ctx, _ptr(b"".join(blobs)), _ptr(b"".join(commitments)), _ptr(b"".join(proofs)),
                                                         ctypes.c_size_t(n), st))
        return list(st[:n])

    def verify_cell_kzg_proof_batches(self, commitments, cell_indices, cells, proofs, batch_offsets):
        """several independent VerifyCellKZGProofBatch verdicts in one call -> list of statuses"""
        N = len(cells)
        nb = len(batch_offsets) - 1
        idx = (ctypes.c_uint64 * max(N, 1))(*cell_indices)
        offs = (ctypes.c_uint64 * (nb + 1))(*batch_offsets)
        res = (ctypes.c_int32 * max(nb, 1))()
        self._check(self.L.kzgb200_verify_cell_kzg_proof_batch(self.ctx, _ptr(b"".join(commitments)) if N else None, idx,
                                                               _ptr(b"".join(cells)) if N else None, _ptr(b"".join(proofs)) if N else None,
                                                               ctypes.c_size_t(N), offs, ctypes.c_size_t(nb), res))
        return list(res[:nb])

    def check_g1_points(self, points48):
        """DeserializeKZGCommitment / DeserializeKZGProof (serialization.go:108-131) validity -> list of statuses"""
        n = len(points48)
        st = (ctypes.c_int32 * max(n, 1))()
        self._check(self.L.kzgb200_check_g1_points(self.ctx, _ptr(b"".join(points48)) if n else None, ctypes.c_size_t(n), st))
        return list(st[:n])

    def check_scalars(self, items, scalars_per_item=1):
        """DeserializeScalar / DeserializeBlob (serialization.go:134-159) validity -> list of statuses"""
        n = len(items)
        st = (ctypes.c_int32 * max(n, 1))()
        self._check(self.L.kzgb200_check_scalars(self.ctx, _ptr(b"".join(items)) if n else None, ctypes.c_size_t(n), ctypes.c_size_t(scalars_per_item), st))
        return list(st[:n])

    # ---- single-item methods, named after the reference's Context methods --------------------
    def blob_to_kzg_commitment(self, blob):
        """Context.BlobToKZGCommitment (prove.go:13-34) -> (status, commitment48)"""
        if len(blob) != BYTES_PER_BLOB:
            return LENGTH_MISMATCH, None
        st, c = self.blob_to_kzg_commitment_batch([blob])[0]
        return st, c

    def compute_kzg_proof(self, blob, z):
        """Context.ComputeKZGProof (prove.go:85-111) -> (status, proof48, y32)"""
        if len(blob) != BYTES_PER_BLOB or len(z) != 32:
            return LENGTH_MISMATCH, None, None
        return self.compute_kzg_proof_batch([blob], [z])[0]

    def compute_blob_kzg_proof(self, blob, commitment):
        """Context.ComputeBlobKZGProof (prove.go:46-77) -> (status, proof48)"""
        if len(blob) != BYTES_PER_BLOB or len(commitment) != 48:
            return LENGTH_MISMATCH, None
        return self.compute_blob_kzg_proof_batch([blob], [commitment])[0]

    def recover_cells_and_kzg_proofs(self, cell_ids, cells):
        """Context.RecoverCellsAndComputeKZGProofs (api_eip7594.go:144-161) -> (status, cells, proofs)"""
        if len(cell_ids) != len(cells):
            return LENGTH_MISMATCH, None, None      # ErrNumCellIDsNotEqualNumCells (api_eip7594.go:94)
        if any(len(c) != BYTES_PER_CELL for c in cells):
            return LENGTH_MISMATCH, None, None
        return self.recover_cells_and_kzg_proofs_batch([list(cell_ids)], [list(cells)])[0]

    def recover_cells(self, cell_ids, cells):
        """Context.RecoverCells (api_eip.go:8-15) -> (status, cells)"""
        if len(cell_ids) != len(cells) or any(len(c) != BYTES_PER_CELL for c in cells):
            return LENGTH_MISMATCH, None
        return self.recover_cells_and_kzg_proofs_batch([list(cell_ids)], [list(cells)], proofs=False)[0]

    def verify_kzg_proof(self, commitment, z, y, proof):
        """Context.VerifyKZGProof (verify.go:12-41) -> status (OK / VERIFY_FAILED / error)"""
        if len(commitment) != 48 or len(z) != 32 or len(y) != 32 or len(proof) != 48:
            return LENGTH_MISMATCH
        return self.verify_kzg_proof_batch([commitment], [z], [y], [proof])[0]

    def verify_blob_kzg_proof(self, blob, commitment, proof):
        """Context.VerifyBlobKZGProof (verify.go:48-82) -> status"""
        if len(blob) != BYTES_PER_BLOB or len(commitment) != 48 or len(proof) != 48:
            return LENGTH_MISMATCH
        return self.verify_blob_kzg_proof_each([blob], [commitment], [proof])[0]

    def verify_blob_kzg_proof_batch(self, blobs, commitments, proofs):
        """Context.VerifyBlobKZGProofBatch (verify.go:88-145): ONE verdict -> status"""
        if not (len(blobs) == len(commitments) == len(proofs)):
            return LENGTH_MISMATCH                         # ErrBatchLengthCheck (verify.go:91-95)
        if any(len(b) != BYTES_PER_BLOB for b in blobs) or any(len(x) != 48 for x in commitments) or any(len(x) != 48 for x in proofs):
            return LENGTH_MISMATCH
        n = len(blobs)
        res = ctypes.c_int32(0)
        self._check(self.L.kzgb200_verify_blob_kzg_proof_batch(self.ctx, _ptr(b"".join(blobs)) if n else None, _ptr(b"".join(commitments)) if n else None,
                                                               _ptr(b"".join(proofs)) if n else None, ctypes.c_size_t(n), ctypes.byref(res)))
        return res.value

    def verify_blob_kzg_proof_batch_par(self, blobs, commitments, proofs):
        """Context.VerifyBlobKZGProofBatchPar (verify.go:152-169): n independent checks, first error wins"""
        if not (len(blobs) == len(commitments) == len(proofs)):
            return LENGTH_MISMATCH
        for st in self.verify_blob_kzg_proof_each(blobs, commitments, proofs):
            if st != OK:
                return st
        return OK

    def verify_cell_kzg_proof_batch(self, commitments, cell_indices, cells, proofs):
        """Context.VerifyCellKZGProofBatch (api_eip7594.go:163-215): one verdict -> status"""
        n = len(commitments)
        if not (n == len(cell_indices) == len(cells) == len(proofs)):
            return LENGTH_MISMATCH                         # ErrBatchLengthCheck (api_eip7594.go:167-171)
        if any(len(x) != 48 for x in commitments) or any(len(x) != BYTES_PER_CELL for x in cells) or any(len(x) != 48 for x in proofs):
            return LENGTH_MISMATCH
        return self.verify_cell_kzg_proof_batches(commitments, cell_indices, cells, proofs, [0, n])[0]

    def compute_cells(self, blob):
        """Context.ComputeCells (api_eip7594.go:12-26) -> (status, cells[128*2048])"""
        if len(blob) != BYTES_PER_BLOB:
            return LENGTH_MISMATCH, None
        return self.compute_cells_and_kzg_proofs_batch([blob], proofs=False)[0]

    def compute_cells_and_kzg_proofs(self, blob):
        """Context.ComputeCellsAndKZGProofs (api_eip7594.go:28-52) -> (status, cells, proofs[128*48])"""
        if len(blob) != BYTES_PER_BLOB:
            return LENGTH_MISMATCH, None, None
        return self.compute_cells_and_kzg_proofs_batch([blob])[0]


class Debug:
    """include/kzgb200_debug.h"""

    def __init__(self):
        self.L = load_library()

    @staticmethod
    def _limbs(vals, n):
        arr = (ctypes.c_uint32 * (len(vals) * n))()
        for i, v in enumerate(vals):
            for k in range(n):
                arr[i * n + k] = (v >> (32 * k)) & 0xffffffff
        return arr

    @staticmethod
    def _ints(arr, cnt, n):
        return [sum(arr[i * n + k] << (32 * k) for k in range(n)) for i in range(cnt)]

    def field_op(self, field, a, b, op):
        n = 12 if field == "fp" else 8
        fn = self.L.kzgb200_dbg_fp_op if field == "fp" else self.L.kzgb200_dbg_fr_op
        A, B = self._limbs(a, n), self._limbs(b, n)
        out = (ctypes.c_uint32 * (len(a) * n))()
        rc = fn(A, B, out, len(a), op)
        if rc:
            raise KzgError(rc, self.L.kzgb200_last_error().decode())
        return self._ints(out, len(a), n)

    def g1_op(self, a48, b48, op):
        n = len(a48)
        out = ctypes.create_string_buffer(48 * n)
        rc = self.L.kzgb200_dbg_g1_op(b"".join(a48), b"".join(b48), out, n, op)
        if rc:
            raise KzgError(rc, self.L.kzgb200_last_error().decode())
        return [out.raw[48 * i:48 * i + 48] for i in range(n)]

    def glv_digits(self, scalars):
        """[(k1_digits[32], k2_digits[32])] of the verifiers' GLV split + signed base-16 recoding"""
        n = len(scalars)
        out = (ctypes.c_int8 * (64 * n))()
        rc = self.L.kzgb200_dbg_glv_digits(self._limbs(scalars, 8), out, n)
        if rc:
            raise KzgError(rc, self.L.kzgb200_last_error().decode())
        return [(list(out[64 * i:64 * i + 32]), list(out[64 * i + 32:64 * i + 64])) for i in range(n)]

    def vmsm(self, ctx, points48, scalars):
        """sum [s_i] P_i through the verifiers' bucket MSM kernels -> compressed point"""
        out = ctypes.create_string_buffer(48)
        rc = self.L.kzgb200_dbg_vmsm(ctx.ctx, b"".join(points48), self._limbs(scalars, 8), len(scalars), out)
        if rc:
            raise KzgError(rc, self.L.kzgb200_last_error().decode())
        return out.raw

    def imad_peak(self, device=0, mode=0):
        """instructions*lanes per second for mad.lo (0), mad.hi (1), mad.wide (2)"""
        v, ms = ctypes.c_double(), ctypes.c_double()
        rc = self.L.kzgb200_bench_imad(device, mode, ctypes.byref(v), ctypes.byref(ms))
        if rc:
            raise KzgError(rc, self.L.kzgb200_last_error().decode())
        return v.value
