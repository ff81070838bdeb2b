"""CPU-side checks of the drop-in boundary: the library loads and exports every symbol the
headers declare (no compute calls: there is no GPU here)."""
import ctypes, os, re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kzgb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    so = os.path.join(ROOT, "go-eth-kzg_b200", "libkzgb200.so")
    assert os.path.exists(so), "run python go-eth-kzg_b200/build.py (or __graft_entry__.build())"
    L = ctypes.CDLL(so)
    names = declared("kzgb200.h") + declared("kzgb200_debug.h")
    assert len(names) >= 18
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_python_mirror_lists_the_same_entry_points():
    import kzgb200
    names = set(declared("kzgb200.h"))
    assert set(kzgb200.EXPORTS) <= names


def test_ctx_new_fails_loudly_without_a_gpu():
    """no CPU fallback: on a box without a CUDA device, context creation must raise"""
    import kzgb200, pytest
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(kzgb200.KzgError) as e:
        kzgb200.Context()
    assert e.value.code == kzgb200.ERR_CUDA
