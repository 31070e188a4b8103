"""Builds libnint.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m nasa_niswan_b200.build

nvcc cross-compiles for sm_100a without a GPU.  The library ships next to this file so the
snapshot taken by gpurun carries it to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["nint_api.cu", "nint_conv_halo.cu", "nint_wgrad.cu", "nint_pointwise.cu"]
HEADERS = ["nint_common.cuh", "nint_kernels.h", "nint_epilogue.cuh", "nint_pair.cuh", os.path.join("..", "..", "include", "nint.h")]
LIB = os.path.join(HERE, "libnint.so")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
