"""Builds the CUDA C-ABI library in-tree for sm_100a (no torch dependency).

    python -m lcp_physics_b200.build        # or __graft_entry__.build()

nvcc cross-compiles without a GPU. The resulting `csrc/liblcpb200.so` is
git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "liblcpb200.so")
STAMP = os.path.join(CSRC, ".liblcpb200.stamp")
SOURCES = ["lcpb200.cu"]
DEPS = ["lcpb200.cu", "lcp_device.cuh", "lcp_solver.cuh", "lcp_assemble.cuh", "../../include/lcpb200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _digest():
    h = hashlib.sha256()
    for d in DEPS:
        with open(os.path.join(CSRC, d), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh():
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == _digest()


def build(force=False, verbose=True):
    if not force and is_fresh():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print("[lcp_physics_b200.build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd, cwd=CSRC)
    with open(STAMP, "w") as f:
        f.write(_digest())
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB)
