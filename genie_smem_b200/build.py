"""Build libgenie_smem.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgenie_smem.so")
SOURCES = ["kernels.cu", "index_host.cpp"]
HEADERS = ["fm_core.cuh", "sweep_logic.cuh", "select_logic.cuh", "host_common.hpp", "../../include/genie_smem.h"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC,-O3,-fno-strict-aliasing", "-shared", "-Xptxas", "-v",
           "-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES] + ["-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libgenie_smem.so")
    with open(os.path.join(HERE, "ptxas.log"), "w") as f:
        f.write(r.stderr)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
