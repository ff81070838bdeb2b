"""ctypes binding of the CPU oracle (oracle/_build/libkzg_oracle.so).  TEST INFRASTRUCTURE."""
import ctypes, os, subprocess, threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_build", "libkzg_oracle.so")
SETUP = os.path.join(ROOT, "go-eth-kzg_b200", "data", "trusted_setup_4096.bin")

OK, VERIFY_FAILED = 0, 1
BLOB, CELLB, NCELLS = 131072, 2048, 128

_lock = threading.Lock()
_state = {}


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    with _lock:
        if "lib" not in _state:
            try:
                build()                      # make: a no-op when the library is newer than its sources
            except Exception:
                if not os.path.exists(SO):
                    raise
            L = ctypes.CDLL(SO)
            L.ko_ctx_new.restype = ctypes.c_void_p
            L.ko_ctx_new.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
            _state["lib"] = L
        return _state["lib"]


def load_setup():
    raw = open(SETUP, "rb").read()
    n = 4096 * 48
    return raw[:n], raw[n:2 * n], raw[2 * n:]


class Oracle:
    """Mirrors the reference's Context method set; each method returns (status, outputs...)."""

    def __init__(self, setup=None):
        L = lib()
        m, l, g2 = setup if setup is not None else load_setup()
        self.L = L
        self.ctx = ctypes.c_void_p(L.ko_ctx_new(m, l, g2, len(g2) // 96))
        assert self.ctx.value, "oracle ctx_new failed"

    def blob_to_kzg_commitment(self, blob):
        if len(blob) != BLOB:
            return 6, None
        out = ctypes.create_string_buffer(48)
        st = self.L.ko_blob_to_kzg_commitment(self.ctx, blob, out)
        return st, out.raw

    def compute_kzg_proof(self, blob, z):
        if len(blob) != BLOB or len(z) != 32:
            return 6, None, None
        p, y = ctypes.create_string_buffer(48), ctypes.create_string_buffer(32)
        st = self.L.ko_compute_kzg_proof(self.ctx, blob, z, p, y)
        return st, p.raw, y.raw

    def compute_blob_kzg_proof(self, blob, commitment):
        if len(blob) != BLOB or len(commitment) != 48:
            return 6, None
        p = ctypes.create_string_buffer(48)
        st = self.L.ko_compute_blob_kzg_proof(self.ctx, blob, commitment, p)
        return st, p.raw

    def verify_kzg_proof(self, c, z, y, proof):
        if len(c) != 48 or len(z) != 32 or len(y) != 32 or len(proof) != 48:
            return 6
        return self.L.ko_verify_kzg_proof(self.ctx, c, z, y, proof)

    def verify_blob_kzg_proof(self, blob, c, proof):
        if len(blob) != BLOB or len(c) != 48 or len(proof) != 48:
            return 6
        return self.L.ko_verify_blob_kzg_proof(self.ctx, blob, c, proof)

    def verify_blob_kzg_proof_batch(self, blobs, cs, proofs):
        if any(len(b) != BLOB for b in blobs) or any(len(c) != 48 for c in cs) or any(len(p) != 48 for p in proofs):
            return 6
        sz = ctypes.c_size_t
        return self.L.ko_verify_blob_kzg_proof_batch(self.ctx, b"".join(blobs), sz(len(blobs)), b"".join(cs), sz(len(cs)),
                                                     b"".join(proofs), sz(len(proofs)))

    def compute_cells(self, blob):
        if len(blob) != BLOB:
            return 6, None
        cells = ctypes.create_string_buffer(NCELLS * CELLB)
        st = self.L.ko_compute_cells(self.ctx, blob, cells)
        return st, cells.raw

    def compute_cells_and_kzg_proofs(self, blob):
        if len(blob) != BLOB:
            return 6, None, None
        cells = ctypes.create_string_buffer(NCELLS * CELLB)
        proofs = ctypes.create_string_buffer(NCELLS * 48)
        st = self.L.ko_compute_cells_and_kzg_proofs(self.ctx, blob, cells, proofs)
        return st, cells.raw, proofs.raw

    def recover_cells_and_kzg_proofs(self, ids, cells):
        if any(len(c) != CELLB for c in cells):
            return 6, None, None
        arr = (ctypes.c_uint64 * max(1, len(ids)))(*ids)
        oc = ctypes.create_string_buffer(NCELLS * CELLB)
        op = ctypes.create_string_buffer(NCELLS * 48)
        sz = ctypes.c_size_t
        st = self.L.ko_recover_cells_and_kzg_proofs(self.ctx, arr, sz(len(ids)), b"".join(cells), sz(len(cells)), oc, op)
        return st, oc.raw, op.raw

    def recover_cells(self, ids, cells):
        if any(len(c) != CELLB for c in cells):
            return 6, None
        arr = (ctypes.c_uint64 * max(1, len(ids)))(*ids)
        oc = ctypes.create_string_buffer(NCELLS * CELLB)
        sz = ctypes.c_size_t
        st = self.L.ko_recover_cells(self.ctx, arr, sz(len(ids)), b"".join(cells), sz(len(cells)), oc)
        return st, oc.raw

    def verify_cell_kzg_proof_batch(self, commitments, cell_indices, cells, proofs):
        if any(len(c) != 48 for c in commitments) or any(len(c) != CELLB for c in cells) or any(len(p) != 48 for p in proofs):
            return 6
        arr = (ctypes.c_uint64 * max(1, len(cell_indices)))(*cell_indices)
        sz = ctypes.c_size_t
        return self.L.ko_verify_cell_kzg_proof_batch(self.ctx, b"".join(commitments), sz(len(commitments)), arr, sz(len(cell_indices)),
                                                     b"".join(cells), sz(len(cells)), b"".join(proofs), sz(len(proofs)))


def rand_blob(seed):
    """bench_test.go:30-46 GetRandBlob"""
    out = ctypes.create_string_buffer(BLOB)
    lib().ko_rand_blob(ctypes.c_int64(seed), out)
    return out.raw


_oracle = {}


R_MOD = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
G1_GEN = bytes.fromhex("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb")


def insecure_setup(secret):
    """(g1_monomial, g1_lagrange, g2_monomial[0..64]) for a KNOWN secret s, as the reference's tests build them
    (internal/kzg/srs_insecure.go:19-58, internal/kzg_multi/srs.go:113-141): [s^i]G1, [L_i(s)]G1 over the natural-order
    4096th roots of unity, [s^i]G2.  The scalars are computed with Python integers; the oracle only multiplies."""
    L = lib()
    n = 4096
    w = pow(7, (R_MOD - 1) // n, R_MOD)
    pows = [pow(secret, i, R_MOD) for i in range(n)]
    zn = (pow(secret, n, R_MOD) - 1) * pow(n, -1, R_MOD) % R_MOD
    lag = [zn * pow(w, i, R_MOD) % R_MOD * pow((secret - pow(w, i, R_MOD)) % R_MOD, -1, R_MOD) % R_MOD for i in range(n)]
    be = lambda xs: b"".join(x.to_bytes(32, "big") for x in xs)
    m = ctypes.create_string_buffer(48 * n); l = ctypes.create_string_buffer(48 * n); g2 = ctypes.create_string_buffer(96 * 65)
    assert L.ko_g1_mul_many(G1_GEN, be(pows), ctypes.c_size_t(n), m) == 0
    assert L.ko_g1_mul_many(G1_GEN, be(lag), ctypes.c_size_t(n), l) == 0
    assert L.ko_g2_mul_many(load_setup()[2][:96], be(pows[:65]), ctypes.c_size_t(65), g2) == 0      # g2_monomial[0] of any setup is the G2 generator
    return m.raw, l.raw, g2.raw


def g1_mul_gen(k):
    """[k]G1 compressed"""
    out = ctypes.create_string_buffer(48)
    assert lib().ko_g1_mul_many(G1_GEN, (k % R_MOD).to_bytes(32, "big"), ctypes.c_size_t(1), out) == 0
    return out.raw


def get_oracle():
    with _lock:
        pass
    if "o" not in _oracle:
        _oracle["o"] = Oracle()
    return _oracle["o"]
