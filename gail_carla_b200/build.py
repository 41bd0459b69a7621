"""Build libgail_carla_b200.so (sm_100a only) in-tree with nvcc.

    python -m gail_carla_b200.build

The shared object is written next to this file (it is git-ignored but travels with the source tree), links the CUDA
runtime statically and resolves cuTensorMapEncodeTiled through cudaGetDriverEntryPoint at run time, so the library
loads (and exports its symbols) on a machine without a GPU driver.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_PATH = os.path.join(HERE, "libgail_carla_b200.so")
SOURCES = ["gc_rollout.cu", "gc_tensor_ops.cu", "gc_umma.cu"]
HEADERS = ["gc_common.cuh", "gc_umma.cuh", "gc_umma_kernel.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr"]


def _flags():
    """GC_UMMA_STATS_BUILD=1 compiles the per-role wait counters of the GEMM kernel in (GC_UMMA_STATS=1 then prints them);
    the production build has no clock reads or counter updates in its loops."""
    return NVCC_FLAGS + (["-DGC_UMMA_STATS_BUILD"] if os.environ.get("GC_UMMA_STATS_BUILD") else [])


def _digest() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    with open(os.path.join(INCLUDE, "gail_carla_b200.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(_flags()).encode())
    return h.hexdigest()


def source_digest() -> str:
    """Digest of the sources next to this file, or "" when they are not there (a binary-only deployment)."""
    try:
        return _digest()
    except OSError:
        return ""


def built_digest() -> str:
    """The digest embedded in the built library (gc_build_digest), "" if there is no library or it predates the export."""
    if not os.path.exists(LIB_PATH):
        return ""
    import ctypes
    try:
        fn = ctypes.CDLL(LIB_PATH).gc_build_digest
    except (OSError, AttributeError):
        return ""
    fn.restype = ctypes.c_char_p
    return fn().decode()


def build_library(force: bool = False, verbose: bool = False) -> str:
    # the digest lives INSIDE the binary (no side-car stamp file that version control could update behind a stale .so)
    dig = _digest()
    if not force and built_digest() == dig:
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *_flags(), f'-DGC_BUILD_DIGEST="{dig}"', "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    subprocess.check_call([nvcc, "-shared", "-o", LIB_PATH, *objs, "-cudart", "static"])
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
