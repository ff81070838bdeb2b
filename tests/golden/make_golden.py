#!/usr/bin/env python3
"""Pack the reference's consensus-spec vectors into a compact fixture.

Run in the authoring container (needs /root/reference):
    python tests/golden/make_golden.py

Reads every  /root/reference/tests/<fn>/kzg-mainnet/<case>/data.y*ml  (the files
consensus_specs_test.go:20-29 globs) and writes
    tests/golden/vectors.json   -- per case: {"fn","name","input","output"} where every hex
                                   string longer than 200 bytes is replaced by {"$ref": idx}
    tests/golden/store.bin.xz   -- the de-duplicated long byte strings, lzma-compressed
    tests/golden/store.idx.json -- [offset, length] per idx into the decompressed store
Also packs the trusted setup (trusted_setup.json) into
    go-eth-kzg_b200/data/trusted_setup_4096.bin  (4096*48 monomial | 4096*48 lagrange | 65*96 G2)
Nothing under /root/reference is needed at test time.
"""
import glob, hashlib, json, lzma, os, sys
import yaml

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

store = bytearray()
index = []          # [offset, length]
by_hash = {}

def intern(b: bytes):
    h = hashlib.sha256(b).digest()
    if h not in by_hash:
        by_hash[h] = len(index)
        index.append([len(store), len(b)])
        store.extend(b)
    return {"$ref": by_hash[h]}

def conv(x):
    if isinstance(x, str) and x.startswith("0x"):
        try:
            b = bytes.fromhex(x[2:])
        except ValueError:
            return x            # odd-length / malformed hex stays a string
        if len(b) > 200:
            return intern(b)
        return x
    if isinstance(x, list):
        return [conv(v) for v in x]
    if isinstance(x, dict):
        return {k: conv(v) for k, v in x.items()}
    return x

def main():
    cases = []
    files = sorted(glob.glob(f"{REF}/tests/*/*/*/*"))
    for f in files:
        parts = f.split("/")
        fn, name = parts[-4], parts[-2]
        with open(f) as fh:
            d = yaml.load(fh, Loader=yaml.CSafeLoader)
        cases.append({"fn": fn, "name": name, "input": conv(d["input"]), "output": conv(d["output"])})
    with open(f"{HERE}/vectors.json", "w") as fh:
        json.dump(cases, fh, separators=(",", ":"))
    with open(f"{HERE}/store.idx.json", "w") as fh:
        json.dump(index, fh, separators=(",", ":"))
    with lzma.open(f"{HERE}/store.bin.xz", "wb", preset=6) as fh:
        fh.write(bytes(store))
    print(f"{len(cases)} cases, {len(index)} stored strings, {len(store)} raw bytes")

    ts = json.load(open(f"{REF}/trusted_setup.json"))
    out = bytearray()
    for k, sz in (("g1_monomial", 48), ("g1_lagrange", 48), ("g2_monomial", 96)):
        for h in ts[k]:
            b = bytes.fromhex(h[2:])
            assert len(b) == sz
            out.extend(b)
    assert len(out) == 4096 * 48 * 2 + 65 * 96, len(out)
    os.makedirs(f"{ROOT}/go-eth-kzg_b200/data", exist_ok=True)
    with open(f"{ROOT}/go-eth-kzg_b200/data/trusted_setup_4096.bin", "wb") as fh:
        fh.write(bytes(out))
    print("trusted setup packed:", len(out), "bytes")

if __name__ == "__main__":
    main()
