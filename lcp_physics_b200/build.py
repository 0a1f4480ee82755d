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
DEPS = ["lcpb200.cu", "lcp_kernels.cu", "lcp_launch.h", "lcp_device.cuh", "lcp_lu.cuh", "lcp_solver.cuh",
        "lcp_assemble.cuh", "../../include/lcpb200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]

# the solver kernels are compiled once per (dtype, residency mode), in parallel
KERNEL_VARIANTS = [(t, m) for t in ("float", "double") for m in (0, 1, 2)]


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
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    objs = []
    for t, m in KERNEL_VARIANTS:
        obj = os.path.join(objdir, "kernels_%s_m%d.o" % (t, m))
        objs.append(obj)
        jobs.append([nvcc] + NVCC_FLAGS + ["-DLCP_T=%s" % t, "-DLCP_MODE=%d" % m, "-c",
                                          os.path.join(CSRC, "lcp_kernels.cu"), "-o", obj])
    api = os.path.join(objdir, "api.o")
    objs.append(api)
    jobs.append([nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, "lcpb200.cu"), "-o", api])
    if verbose:
        for j in jobs:
            print("[lcp_physics_b200.build]", " ".join(j), flush=True)
    procs = [subprocess.Popen(j, cwd=CSRC) for j in jobs]
    rcs = [p.wait() for p in procs]
    if any(rcs):
        raise RuntimeError("nvcc failed (exit codes %s)" % rcs)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs
    if verbose:
        print("[lcp_physics_b200.build]", " ".join(link), flush=True)
    subprocess.check_call(link, cwd=CSRC)
    with open(STAMP, "w") as f:
        f.write(_digest())
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB)
