"""Builds librelp_gpu.so (CUDA engine + C++ host driver) in-tree for sm_100a.

The K1 variants (k1_variants.cu, one translation unit per limb width) dominate the compile time, so the
translation units are compiled in parallel and linked afterwards."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "librelp_gpu.so")
OBJ = os.path.join(HERE, "_obj")
CSRC = os.path.join(HERE, "csrc")
K1_WIDTHS = (16, 14, 12, 10, 8, 4, 2, 1)          # widest first: it is the longest job
SOURCES = [
    os.path.join(CSRC, "relp_gpu.cu"),
    os.path.join(CSRC, "k1_variants.cu"),
    os.path.join(CSRC, "host", "relp_host.cpp"),
]
HEADERS = [
    os.path.join(CSRC, "bigint.cuh"),
    os.path.join(CSRC, "mp32.cuh"),
    os.path.join(CSRC, "engine.cuh"),
    os.path.join(CSRC, "kernels.cuh"),
    os.path.join(CSRC, "k1_update.cuh"),
    os.path.join(ROOT, "include", "relp_gpu.h"),
    os.path.join(ROOT, "include", "relp_gpu_test.h"),
    os.path.join(ROOT, "include", "relp_host.h"),
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
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
    os.makedirs(OBJ, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []
    jobs = []
    for L in K1_WIDTHS:
        obj = os.path.join(OBJ, f"k1_L{L}.o")
        jobs.append((obj, [nvcc] + NVCC_FLAGS + extra + ["--split-compile", "2", f"-DRG_K1_L={L}", "-c",
                                                          os.path.join(CSRC, "k1_variants.cu"), "-o", obj]))
    obj = os.path.join(OBJ, "relp_gpu.o")
    jobs.append((obj, [nvcc] + NVCC_FLAGS + extra + ["--split-compile", "2", "-c", os.path.join(CSRC, "relp_gpu.cu"),
                                                      "-o", obj]))
    obj = os.path.join(OBJ, "relp_host.o")
    jobs.append((obj, [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, "host", "relp_host.cpp"), "-o", obj]))

    def run(job):
        print("[relp_b200.build]", " ".join(job[1]), file=sys.stderr)
        subprocess.run(job[1], check=True)
        return job[0]

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(run, jobs))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB] + objs
    print("[relp_b200.build]", " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
