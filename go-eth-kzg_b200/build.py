"""Builds libkzgb200.so (CUDA, sm_100a) in-tree.  nvcc cross-compiles without a GPU."""
import glob, os, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libkzgb200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    root = os.path.dirname(HERE)
    return _sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(root, "include", "*.h"))


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(f) > t for f in _deps())


# per-file extra flags.  kzgb200_verify.cu (host side of the verifiers + latency-bound preparation kernels): ptxas -O1
# as a guard against a miscompilation met there (header of that file)
PER_FILE = {"kzgb200_verify.cu": os.environ.get("KZGB200_VERIFY_FLAGS", "-Xptxas -O1").split()}


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("KZGB200_NVCC_EXTRA", "").split()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        flags = [f for f in NVCC_FLAGS if f != "-shared"]
        cmd = [nvcc] + flags + PER_FILE.get(os.path.basename(src), []) + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        print("[kzgb200] " + " ".join(cmd), file=sys.stderr)
        procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    cmd = [nvcc, "-Wno-deprecated-gpu-targets", "-shared", "-o", SO] + objs
    print("[kzgb200] " + " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
