"""Trusted-setup ingestion.  The JSON parser is host code and runs without a GPU; the
context-from-JSON and CheckTrustedSetupIsWellFormed paths need one."""
import ctypes, json, random
import pytest
import oracle_lib


def setup_json_text(m=None, l=None, g2=None, prefix="0x", extra=None):
    """a JSONTrustedSetup document (trusted_setup.go:23-27) made from the packed mainnet setup"""
    sm, sl, sg = oracle_lib.load_setup()
    m, l, g2 = m or sm, l or sl, g2 or sg
    doc = {"g2_monomial": [prefix + g2[96 * i:96 * i + 96].hex() for i in range(len(g2) // 96)],
           "g1_lagrange": [prefix + l[48 * i:48 * i + 48].hex() for i in range(len(l) // 48)],
           "g1_monomial": [prefix + m[48 * i:48 * i + 48].hex() for i in range(len(m) // 48)]}
    if extra:
        doc.update(extra)
    return json.dumps(doc, indent=1)


def test_json_parser_round_trip_and_errors():
    import kzgb200
    m, l, g2 = oracle_lib.load_setup()
    assert kzgb200.parse_trusted_setup_json(setup_json_text()) == (m, l, g2)
    # no 0x prefix, upper case, unknown keys of every JSON shape, compact separators
    doc = json.loads(setup_json_text(prefix="", extra={"comment": "x ] } \" y", "n": 4096, "nested": {"a": [1, {"b": "]"}]}, "flag": True}))
    doc["g1_monomial"][7] = "0X" + doc["g1_monomial"][7].upper()
    assert kzgb200.parse_trusted_setup_json(json.dumps(doc, separators=(",", ":"))) == (m, l, g2)
    bad_docs = []
    d = json.loads(setup_json_text()); d["g1_lagrange"][5] = d["g1_lagrange"][5][:-2]; bad_docs.append(d)            # 47 bytes
    d = json.loads(setup_json_text()); d["g1_monomial"][9] = "0xzz" + d["g1_monomial"][9][4:]; bad_docs.append(d)     # not hex
    d = json.loads(setup_json_text()); d["g1_monomial"] = d["g1_monomial"][:-1]; bad_docs.append(d)                  # 4095 points
    d = json.loads(setup_json_text()); del d["g2_monomial"]; bad_docs.append(d)                                      # missing key
    d = json.loads(setup_json_text()); d["g2_monomial"][0] = d["g2_monomial"][0][:-2]; bad_docs.append(d)            # short G2
    for d in bad_docs:
        with pytest.raises(kzgb200.KzgError) as e:
            kzgb200.parse_trusted_setup_json(json.dumps(d))
        assert e.value.code == 12                                                                                   # KZGB200_ERR_SETUP
    # a key that appears twice: the last value wins (encoding/json) and nothing is written past the 4096-point buffers
    # (the parser used to append, so two copies of a key overflowed the caller's buffer)
    text = setup_json_text()
    dup = json.loads(text)
    junk = ["0x" + bytes([i % 251] * 48).hex() for i in range(4096)]
    for key in ("g1_monomial", "g1_lagrange"):
        doubled = '{"%s": %s, %s' % (key, json.dumps(junk), text.lstrip()[1:])
        assert kzgb200.parse_trusted_setup_json(doubled) == (m, l, g2)
    doubled = '{"g2_monomial": %s, %s' % (json.dumps(dup["g2_monomial"] * 3), text.lstrip()[1:])
    assert kzgb200.parse_trusted_setup_json(doubled) == (m, l, g2)
    short_last = text.rstrip()[:-1] + ', "g1_monomial": %s}' % json.dumps(junk[:100])            # last copy has 100 points -> rejected
    with pytest.raises(kzgb200.KzgError) as e:
        kzgb200.parse_trusted_setup_json(short_last)
    assert e.value.code == 12
    for text in ("", "[]", "{", '{"g1_monomial": [', setup_json_text()[:-3]):
        with pytest.raises(kzgb200.KzgError):
            kzgb200.parse_trusted_setup_json(text)


@pytest.mark.gpu
def test_context_from_json_matches_packed_setup():
    import kzgb200
    blob = oracle_lib.rand_blob(5 << 20)
    exp = oracle_lib.get_oracle().blob_to_kzg_commitment(blob)[1]
    c = kzgb200.Context(commit_window=8, fk20_window=8, setup_json=setup_json_text())
    try:
        assert c.blob_to_kzg_commitment(blob) == (0, exp)
        st, cells, proofs = c.compute_cells_and_kzg_proofs(blob)
        assert st == 0 and c.verify_cell_kzg_proof_batch([exp] * 128, list(range(128)), [cells[2048 * i:2048 * i + 2048] for i in range(128)],
                                                          [proofs[48 * i:48 * i + 48] for i in range(128)]) == 0
    finally:
        c.close()
    with pytest.raises(kzgb200.KzgError):                       # fewer than 65 G2 points (api.go:93,115)
        kzgb200.Context(commit_window=8, fk20_window=8, setup_json=setup_json_text(g2=oracle_lib.load_setup()[2][:96 * 64]))


@pytest.mark.gpu
def test_g2_group_law_self_consistency():
    """decode, on-curve, 2Q, 3Q, 4Q two ways, the addition's doubling and cancellation branches, [r]Q == O twice"""
    import kzgb200
    L = kzgb200.load_library()
    g2 = oracle_lib.load_setup()[2]
    for i in (0, 1, 2, 64):
        m = ctypes.c_int()
        assert L.kzgb200_dbg_g2_selftest(g2[96 * i:96 * i + 96], ctypes.byref(m)) == 0
        assert m.value == 0x1ff, (i, bin(m.value))


def _g2_oracle(p96, subgroup=1):
    return oracle_lib.lib().ko_g2_decompress(p96, subgroup)


@pytest.mark.gpu
def test_check_trusted_setup_is_well_formed():
    """trusted_setup.go:45-83: order Lagrange, monomial, G2; first failure wins; G2 decisions equal the oracle's [r]Q test"""
    import kzgb200
    m, l, g2 = oracle_lib.load_setup()
    assert kzgb200.check_trusted_setup(l, m, g2) == (0, 0)
    P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    rng = random.Random(21)
    # a G1 point on the curve but outside the subgroup (random x; the cofactor makes subgroup members rare)
    L = oracle_lib.lib()
    while True:
        b = bytearray(rng.randrange(P).to_bytes(48, "big")); b[0] |= 0x80
        if L.ko_g1_decompress(bytes(b), ctypes.create_string_buffer(96), 1) == 5:
            break
    bad_g1 = bytes(b)
    lm = bytearray(m); lm[48 * 100:48 * 101] = bad_g1
    assert kzgb200.check_trusted_setup(l, bytes(lm), g2) == (5, 4096 + 100)
    ll = bytearray(l); ll[48 * 7] ^= 0x40                                      # infinity flag on a non-zero body
    assert kzgb200.check_trusted_setup(bytes(ll), bytes(lm), g2) == (3, 7)      # the Lagrange basis is checked first
    # G2: random x in Fp2 -> bad encoding / not on curve / on the twist but outside G2; plus valid points and infinity
    pts = []
    for _ in range(12):
        c1, c0 = rng.randrange(P), rng.randrange(P)
        b = bytearray(c1.to_bytes(48, "big") + c0.to_bytes(48, "big")); b[0] |= 0x80 | (0x20 if rng.random() < 0.5 else 0)
        pts.append(bytes(b))
    pts += [g2[:96], g2[96:192], bytes([0xc0]) + bytes(95), bytes([0xc0]) + bytes(94) + b"\x01", bytes([0x20]) + g2[1:96],
            bytes([g2[0] ^ 0x20]) + g2[1:96]]                                    # the last one: -G2, still in the subgroup
    exp = [_g2_oracle(p) for p in pts]
    assert 5 in exp and 4 in exp and 0 in exp and 3 in exp
    for p, e in zip(pts, exp):
        got = kzgb200.check_trusted_setup(b"", b"", p)
        assert got == (e, 0), (p.hex()[:16], got, e)
    # a bad G2 point is reported after clean G1 bases, with its index in the concatenation
    first_bad = next(p for p, e in zip(pts, exp) if e == 5)
    assert kzgb200.check_trusted_setup(l, m, g2[:96 * 3] + first_bad + g2[96 * 4:]) == (5, 8192 + 3)


def test_json_parser_never_crashes_on_garbage():
    """the parser reads untrusted text: truncations, byte flips and random insertions must end in a clean error or a
    well-formed result, never in a crash (the process surviving this loop is the assertion)"""
    import kzgb200
    rng = random.Random(77)
    m, l, g2 = oracle_lib.load_setup()
    # a small but structurally complete document keeps the loop fast: the parser insists on 4096 points, so every
    # mutation of this text must be rejected cleanly; full-size mutations are tried a few times
    small = json.dumps({"g1_monomial": ["0x" + m[:48].hex()] * 3, "g1_lagrange": ["0x" + l[:48].hex()] * 3, "g2_monomial": ["0x" + g2[:96].hex()]})
    full = setup_json_text()
    alphabet = b'{}[]",:\\0xX19afAF \n\t\x00\xff'
    for trial in range(600):
        base = full if trial % 100 == 0 else small
        b = bytearray(base.encode())
        for _ in range(rng.randrange(1, 6)):
            op, pos = rng.randrange(4), rng.randrange(len(b))
            if op == 0: b[pos] = rng.choice(alphabet)
            elif op == 1: del b[pos:pos + rng.randrange(1, 40)]
            elif op == 2: b[pos:pos] = bytes(rng.choice(alphabet) for _ in range(rng.randrange(1, 8)))
            else: b = b[:pos]
            if not b:
                break
        try:
            out = kzgb200.parse_trusted_setup_json(bytes(b))
            assert len(out[0]) == 4096 * 48 and len(out[1]) == 4096 * 48 and len(out[2]) % 96 == 0
        except kzgb200.KzgError as e:
            assert e.code in (11, 12)
