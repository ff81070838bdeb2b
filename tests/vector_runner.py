"""Runs the 311 consensus-spec vectors against any object with the Context-like method set
(the oracle binding or the CUDA binding).  Mirrors consensus_specs_test.go: outputs must be
byte-equal; verifiers map {OK, VERIFY_FAILED, anything else} onto {true, false, null}."""
from golden_util import cases, resolve, BadHex

OK, VERIFY_FAILED = 0, 1


def _r(x):
    return resolve(x)


def run_case(impl, c):
    """returns (got, expected) in a normalised comparable form"""
    fn, inp, out = c["fn"], c["input"], c["output"]
    try:
        if fn == "blob_to_kzg_commitment":
            st, cm = impl.blob_to_kzg_commitment(_r(inp["blob"]))
            got = cm if st == OK else None
            exp = _r(out) if out is not None else None
        elif fn == "compute_kzg_proof":
            st, p, y = impl.compute_kzg_proof(_r(inp["blob"]), _r(inp["z"]))
            got = (p, y) if st == OK else None
            exp = (_r(out[0]), _r(out[1])) if out is not None else None
        elif fn == "compute_blob_kzg_proof":
            st, p = impl.compute_blob_kzg_proof(_r(inp["blob"]), _r(inp["commitment"]))
            got = p if st == OK else None
            exp = _r(out) if out is not None else None
        elif fn == "verify_kzg_proof":
            st = impl.verify_kzg_proof(_r(inp["commitment"]), _r(inp["z"]), _r(inp["y"]), _r(inp["proof"]))
            got = True if st == OK else False if st == VERIFY_FAILED else None
            exp = out
        elif fn == "verify_blob_kzg_proof":
            st = impl.verify_blob_kzg_proof(_r(inp["blob"]), _r(inp["commitment"]), _r(inp["proof"]))
            got = True if st == OK else False if st == VERIFY_FAILED else None
            exp = out
        elif fn == "verify_blob_kzg_proof_batch":
            st = impl.verify_blob_kzg_proof_batch([_r(b) for b in inp["blobs"]], [_r(b) for b in inp["commitments"]],
                                                  [_r(b) for b in inp["proofs"]])
            got = True if st == OK else False if st == VERIFY_FAILED else None
            exp = out
        elif fn == "compute_cells_and_kzg_proofs":
            st, cells, proofs = impl.compute_cells_and_kzg_proofs(_r(inp["blob"]))
            got = (cells, proofs) if st == OK else None
            exp = (b"".join(_r(x) for x in out[0]), b"".join(_r(x) for x in out[1])) if out is not None else None
        elif fn == "recover_cells_and_kzg_proofs":
            st, cells, proofs = impl.recover_cells_and_kzg_proofs(list(inp["cell_indices"]), [_r(x) for x in inp["cells"]])
            got = (cells, proofs) if st == OK else None
            exp = (b"".join(_r(x) for x in out[0]), b"".join(_r(x) for x in out[1])) if out is not None else None
        elif fn == "verify_cell_kzg_proof_batch":
            st = impl.verify_cell_kzg_proof_batch([_r(x) for x in inp["commitments"]], list(inp["cell_indices"]),
                                                  [_r(x) for x in inp["cells"]], [_r(x) for x in inp["proofs"]])
            got = True if st == OK else False if st == VERIFY_FAILED else None
            exp = out
        else:
            raise KeyError(fn)
    except BadHex:
        got = None
        exp = out if fn.startswith("verify") else (None if out is None else "non-null")
    return got, exp


ALL_FNS = ["blob_to_kzg_commitment", "compute_kzg_proof", "compute_blob_kzg_proof", "verify_kzg_proof",
           "verify_blob_kzg_proof", "verify_blob_kzg_proof_batch", "compute_cells_and_kzg_proofs",
           "recover_cells_and_kzg_proofs", "verify_cell_kzg_proof_batch"]
