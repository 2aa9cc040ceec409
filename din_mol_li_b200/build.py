"""Builds libdml.so (sm_100a) in-tree with nvcc.  No torch involved: the product is a plain C-ABI library."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdml.so")
SRCS = ["dml.cu", "dana_host.cpp"]
import glob
DEPS = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.cpp")) +
              glob.glob(os.path.join(HERE, "..", "include", "*.h")) + [os.path.join(HERE, "..", "tools", "dana_host.cpp")])
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v", "-ldl"]


def build(force=False, verbose=False):
    deps = DEPS
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
    # C++ stand-in for the reference's main program, linked against the library (tools/dana_host.cpp)
    exe = os.path.join(HERE, "dana_b200")
    cmd = [NVCC, "-O2", "-std=c++17", "-o", exe, os.path.join(HERE, "..", "tools", "dana_host.cpp"), "-L" + HERE, "-ldml",
           "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building dana_b200")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
