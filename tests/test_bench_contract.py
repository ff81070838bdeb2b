"""bench.py contract, the part that runs without a GPU: the reference arm prints exactly one JSON line with the keys
the driver reads (the GPU arm's line is checked by the driver itself and archived under profiles/)."""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "commit", "--steps", "9", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "BlobToKZGCommitment blobs/s" and d["unit"] == "blobs/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 9 and d["warmup"] == 3
    assert d["e2e"] == {"value": d["value"], "unit": "blobs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "blobs" in cb["sample"]
    assert d["config"]["workload"].startswith("EIP-4844 BlobToKZGCommitment")


def test_reference_arm_covers_the_verifier_workloads():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "verify_cells", "--steps", "9", "--warmup", "3"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "cells/s" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
