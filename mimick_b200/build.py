"""Compile libmimick_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python mimick_b200/build.py [--force] [-v]

The shared library links the CUDA runtime statically and zlib dynamically; it has no Python or
torch dependency (the C ABI in include/mimick_b200.h is the product boundary).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(HERE, "libmimick_b200.so")
OBJ = os.path.join(HERE, "_build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -fmad is irrelevant for the variate transforms (they use __dmul_rn/__dadd_rn), but host code in
# the same TUs must not contract either.
HOST_FLAGS = "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall,-Wno-unused-function"
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", HOST_FLAGS, "-I", INCLUDE, "-I", CSRC]
COMMON += os.environ.get("MMK_NVCC_FLAGS", "").split()  # experiment switches (-DMMK_EMIT_ILP=4 ...); the defaults are the product

SOURCES = ["mmk_device.cu", "mmk_host.cpp", "mmk_io.cpp"]
HEADERS = ["mmk_internal.h", "mmk_sampling.cuh", "mmk_host.h", "mmk_emit.cuh", "mmk_bgzf.cuh", "mmk_bgzf_table.h", "mmk_place.cuh", "mmk_plan.cuh"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmimick_b200.so cannot be built (there is no CPU fallback)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [
        os.path.join(INCLUDE, "mimick_b200.h"),
        os.path.join(INCLUDE, "mmk_rng.h"),
        os.path.abspath(__file__),
    ]
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + ARCH + COMMON + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(r.stderr)
    if force or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lz", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
