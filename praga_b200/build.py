"""Build libpraga_b200.so (hand-written CUDA for sm_100a + the C ABI of include/praga_b200.h) in-tree with nvcc.

The shared library has no torch dependency: PyTorch is only used by bench.py for torch.distributed plumbing.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpraga_b200.so")
SOURCES = ["engine.cu", "idw_kernels.cu", "peaks.cu"]
HEADERS = ["common.cuh", os.path.join("..", "..", "include", "praga_b200.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: praga_b200 needs the CUDA toolkit to build (there is no CPU fallback)")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into praga_b200/libpraga_b200.so. Returns the library path."""
    if not force and not is_stale():
        return LIB
    cmd = [
        nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
        "-shared", "-Xcompiler", "-fPIC,-fopenmp,-O2", "-ccbin", "/usr/bin/g++",
        "-Xptxas", "-v" if verbose else "-O3",
        "-o", LIB,
    ] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lgomp"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libpraga_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
