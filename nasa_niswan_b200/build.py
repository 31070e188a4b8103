"""Builds libnint.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m nasa_niswan_b200.build [--force] [-v] [--knobs]

nvcc cross-compiles for sm_100a without a GPU.  The library ships next to this file so the
snapshot taken by gpurun carries it to the GPU box.  Every translation unit is compiled to an
object file (in parallel, only when it or a header is newer) and the objects are linked into the
shared library.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SOURCES = ["nint_api.cu", "nint_conv_halo.cu", "nint_wgrad.cu", "nint_pointwise.cu", "nint_dp.cu"]
HEADERS = ["nint_common.cuh", "nint_kernels.h", "nint_epilogue.cuh", "nint_pair.cuh", os.path.join("..", "..", "include", "nint.h")]
LIB = os.path.join(HERE, "libnint.so")
LIB_KNOBS = os.path.join(HERE, "libnint_knobs.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _mtime(path):
    return os.path.getmtime(path) if os.path.exists(path) else 0.0


def _stale(lib=None) -> bool:
    t = _mtime(lib or LIB)
    return t == 0.0 or any(_mtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, knobs: bool = False) -> str:
    """knobs: the experiment build (-DNINT_KNOBS=1 -> libnint_knobs.so): the kernels honour NINT_DEBUG_FLAGS (skip the
    epilogue's memory work, no MMA issue, timeline stamps, ...); the product build compiles those branches out"""
    lib = LIB_KNOBS if knobs else LIB
    if not force and not _stale(lib):
        return lib
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    obj_dir = OBJ + ("_knobs" if knobs else "")
    os.makedirs(obj_dir, exist_ok=True)
    hdr_time = max(_mtime(os.path.join(CSRC, h)) for h in HEADERS)
    flags = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"] + (["-DNINT_KNOBS=1"] if knobs else [])
    if verbose:
        flags += ["-Xptxas", "-v"]

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if force or _mtime(obj) < max(_mtime(path), hdr_time):
            subprocess.run([nvcc] + flags + ["-c", path, "-o", obj], check=True)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    subprocess.run([nvcc] + ARCH + ["-shared", "-o", lib] + objs, check=True)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, knobs="--knobs" in sys.argv))
