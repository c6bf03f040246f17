"""In-tree build of the C-ABI CUDA library (plain nvcc, no torch headers).

    python -m aas_enhancement_b200.build          # -> aas_enhancement_b200/libaas_lmfb.so
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "lmfb_kernels.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "lmfb_core.cuh"), os.path.join(HERE, "csrc", "fft_codelets.cuh"), os.path.join(HERE, "csrc", "mel_band.hpp"),
        os.path.join(os.path.dirname(HERE), "include", "aas_lmfb.h")]
LIB = os.path.join(HERE, "libaas_lmfb.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "--split-compile", "0"]      # (the kernels are optimised in parallel)


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the LMFB front-end has no non-CUDA path")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    extra = os.environ.get("AAS_LMFB_NVCC_EXTRA", "").split()      # e.g. -DLMFB_ONLY_W3 for quick experiments
    cmd = [find_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + [SRC, "-o", LIB]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
