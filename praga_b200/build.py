"""Build libpraga_b200.so (hand-written CUDA for sm_100a + the C ABI of include/praga_b200.h) in-tree with nvcc.

The shared library has no torch dependency: PyTorch is only used by bench.py for torch.distributed plumbing.
Each .cu is compiled to an object with its own flags, then linked:
  the TUs on the bit-exact station-space / per-cell fit paths (local_kernels.cu, pre_kernels.cu, classic_kernels.cu,
  station_space.cu, td_kernels.cu, shepard_kernels.cu, glocal.cu, raster_ops.cu) are built with -fmad=false: they follow
  the reference's FP32 / FP64 operation order (no FMA contraction, like the reference's x86-64 build) so that
  Levenberg-Marquardt takes the same accept / reject path and every decision falls the same way.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpraga_b200.so")
SOURCES = {
    "engine.cu": [],
    "idw_kernels.cu": [],
    "peaks.cu": [],
    "local_kernels.cu": ["-fmad=false"],
    "local_pipeline.cu": ["-fmad=false", "-DPB_LM_UNROLL=3"],
    "pre_kernels.cu": ["-fmad=false"],
    "td_kernels.cu": ["-fmad=false"],
    "classic_kernels.cu": ["-fmad=false"],
    "shepard_kernels.cu": ["-fmad=false"],
    "station_space.cu": ["-fmad=false"],
    "glocal.cu": ["-fmad=false"],
    "raster_ops.cu": ["-fmad=false"],
    "radiation.cu": ["-fmad=false"],
    "raster_io.cu": [],
    "kriging.cu": [],
    "kriging_tc.cu": [],
    "variogram.cu": [],
}
HEADERS = ["common.cuh", "epilogue.cuh", "context.cuh", "local.cuh", "local_device.cuh", "nvtx.cuh", "td_device.cuh", "kh_search.cuh", "pre.cuh", "multiple_detrend.cuh", "detrend_device.cuh", "kriging.cuh", "classic.cuh", "shepard.cuh", os.path.join("..", "..", "include", "praga_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: praga_b200 needs the CUDA toolkit to build (there is no CPU fallback)")


def _deps():
    return [os.path.join(CSRC, s) for s in list(SOURCES) + HEADERS] + [os.path.abspath(__file__)]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps() if os.path.exists(d))


def _run(cmd, verbose):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd[-3:]))


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into praga_b200/libpraga_b200.so. Returns the library path."""
    if not force and not is_stale():
        return LIB
    nvcc = nvcc_path()
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS if os.path.exists(os.path.join(CSRC, h)))
    hdr_t = max(hdr_t, os.path.getmtime(os.path.abspath(__file__)))
    objs, procs = [], []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if not force and os.path.exists(o) and os.path.getmtime(o) > max(os.path.getmtime(s), hdr_t):
            continue
        cmd = [nvcc] + ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fopenmp,-O2", "-ccbin", "/usr/bin/g++"]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        if os.environ.get("PB_DEBUG"):
            cmd += ["-DPB_DEBUG_ASSERTS"]          # device-side capacity asserts (common.cuh PB_DEV_ASSERT)
        cmd += extra + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libpraga_b200.so")
    _run([nvcc] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs + ["-lgomp"], verbose)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
