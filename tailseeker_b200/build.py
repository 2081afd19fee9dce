"""Builds libtailseeker_b200.so IN-TREE with nvcc for sm_100a (cross-compiles without a GPU).

The library is the product: CUDA kernels + the extern "C" boundary of
include/tailseeker_b200.h.  -fmad=false is deliberate: the reference is compiled without
FMA contraction and parity of the packed scores needs separately rounded float operations.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtailseeker_b200.so")
SOURCES = ["tsk_kernels.cu", "tsk_score.cu", "tsk_capi.cu", "tsk_ruler.cu", "tsk_fit.cu", "tsk_control.cu",
           "tsk_dedup.cu", "tsk_encode.cu", "tsk_merge.cu", "tsk_ingest.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false", "--shared", "-Xcompiler", "-fPIC", "-cudart", "static", "-ldl"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "tailseeker_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    extra = os.environ.get("TSK_NVCC_EXTRA", "").split()      # experiments only (e.g. -DNAME=1)
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libtailseeker_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
