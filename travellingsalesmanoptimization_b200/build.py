"""Build libtspcuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtspcuda.so")
SOURCES = [os.path.join(CSRC, "tspcuda.cu")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("kernels.cuh", "metric.cuh", "pruned.cuh", "matrix.cuh", "mileage.cuh", "patching.cuh", "mgpu.cuh")] + [
    os.path.join(HERE, "..", "include", "tspcuda.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]
# TSPCUDA_FAST_BUILD=1: parallel back end (development only -- measured: -split-compile changes the code generated for
# descend_kernel, 2344 instead of 2240 instructions, 3.6 % fewer pair evaluations per second).  Not part of the source hash.
SPEED_FLAGS = ["-split-compile", str(min(8, os.cpu_count() or 1))] if os.environ.get("TSPCUDA_FAST_BUILD") == "1" else []


STAMP = LIB + ".srchash"      # content hash of the sources the library was built from (mtimes do not survive a snapshot copy)


def _source_hash():
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in DEPS:
        if os.path.exists(d):
            with open(d, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _source_hash()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + SPEED_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.check_call(cmd, cwd=CSRC)
    with open(STAMP, "w") as f:
        f.write(_source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
