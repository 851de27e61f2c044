"""Builds libries_b200.so (CUDA kernels + host driver + C ABI) and the `ries` binary, in-tree.

nvcc cross-compiles sm_100a without a GPU.  -fmad=false is REQUIRED: the reference dedupes on `==`
of doubles, so no multiply-add may be contracted unless the source says fma() (ries.c:7651 rejects
-ffast-math for the same reason).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libries_b200.so")
BIN = os.path.join(HERE, "bin", "ries")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fvisibility=hidden", "-I", CSRC,
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def sources():
    out = []
    for root, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cc", ".h", ".c", ".inc")):
                out.append(os.path.join(root, f))
    out.append(os.path.join(HERE, "..", "include", "ries_b200.h"))
    return out


def build(verbose=False, force=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if force or _newer(LIB, sources()):
        srcs = [os.path.join(CSRC, "engine.cu"), os.path.join(CSRC, "capi.cc")]
        srcs += sorted(os.path.join(CSRC, "host", f) for f in os.listdir(os.path.join(CSRC, "host")) if f.endswith(".cc"))
        cmd = [nvcc] + NVCC_FLAGS + ["-shared", "-Xcompiler", "-fvisibility=default", "-o", LIB] + srcs
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    if force or _newer(BIN, [LIB, os.path.join(CSRC, "bin", "ries_main.c")]):
        os.makedirs(os.path.dirname(BIN), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-o", BIN, os.path.join(CSRC, "bin", "ries_main.c"),
                               "-L", HERE, "-lries_b200", "-Wl,-rpath,$ORIGIN/.."])
    return LIB


if __name__ == "__main__":
    build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(LIB)
