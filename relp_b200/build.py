"""Builds librelp_gpu.so (CUDA engine + C++ host driver) in-tree for sm_100a."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "librelp_gpu.so")
SOURCES = [
    os.path.join(HERE, "csrc", "relp_gpu.cu"),
    os.path.join(HERE, "csrc", "host", "relp_host.cpp"),
]
HEADERS = [
    os.path.join(HERE, "csrc", "bigint.cuh"),
    os.path.join(HERE, "csrc", "mp32.cuh"),
    os.path.join(HERE, "csrc", "engine.cuh"),
    os.path.join(HERE, "csrc", "kernels.cuh"),
    os.path.join(ROOT, "include", "relp_gpu.h"),
    os.path.join(ROOT, "include", "relp_host.h"),
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--split-compile", "0",     # the kernel templates expand to ~300 instantiations: optimise / assemble them in parallel
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    print("[relp_b200.build]", " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
