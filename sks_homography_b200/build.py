"""Builds libsks_cuda.so (the product) in-tree with nvcc for sm_100a.

No JIT, no torch extension machinery: the library is plain CUDA C++ behind a
C ABI, so an explicit nvcc command is the whole build.  The .so is git-ignored
but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsks_cuda.so")
SOURCES = ["capi.cu", "host_api.cu", "multi.cu"]
HEADERS = ["strict.cuh", "solvers.cuh", "ptx.cuh", "stream_kernels.cuh", "synth.cuh", "ransac.cuh", "peer.cuh", "mrg32k3a.cuh", "warp.cuh", "refit.cuh",
           os.path.join("..", "..", "include", "sks_cuda.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # bit-exact parity with the reference's -ffp-contract=off build: never fuse,
    # IEEE division / sqrt, keep denormals (see csrc/strict.cuh)
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libsks_cuda cannot be built (there is no CPU fallback)")
    return exe


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libsks_cuda.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
