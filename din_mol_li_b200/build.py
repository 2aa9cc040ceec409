"""Builds libdml.so (sm_100a) in-tree with nvcc.  No torch involved: the product is a plain C-ABI library."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdml.so")
SRCS = ["dml.cu", "dana_host.cpp"]
DEPS = ["dml.cu", "dana_host.cpp", os.path.join("..", "..", "include", "dml_host.h"), "dml_kernels.cuh", "dml_device.cuh", "dml_gcmc.cuh", os.path.join("..", "..", "include", "dml.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def build(force=False, verbose=False):
    deps = [os.path.join(CSRC, d) for d in DEPS]
    if (not force) and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    cmd = [NVCC] + FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SRCS]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    open(os.path.join(HERE, "build.log"), "w").write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libdml.so")
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
