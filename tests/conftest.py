import os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    """The built artefacts are git-ignored: in a fresh checkout (no libkzgb200.so at all) build the library before the
    first test needs it (nvcc cross-compiles without a GPU).  An existing library is left alone -- keeping it current is
    __graft_entry__.build()'s job.  The oracle builds itself on first use (oracle_lib)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("kzgb200_build", os.path.join(ROOT, "go-eth-kzg_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if os.path.exists(mod.SO):
        return
    try:
        mod.build()
    except Exception as e:          # tests that need the library will say so themselves
        print("[conftest] libkzgb200.so build failed:", e, file=sys.stderr)
