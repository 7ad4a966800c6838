"""Builds rdcfes_b200/librdcgpu.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m rdcfes_b200.build [--force]

nvcc cross-compiles without a GPU.  FMA contraction stays on; the arithmetic that feeds discrete decisions
uses __dmul_rn/__dadd_rn explicitly (csrc/models.cuh "Rounding discipline").  -lineinfo keeps ncu's source
page usable.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "librdcgpu.so")
SOURCES = ["api.cu", "assemble.cu", "solver.cu", "solid.cu", "p2p.cu", "reduce.cu", "setup.cpp", "comm.cpp"]
HEADERS = ["rdc_internal.h", "models.cuh", "asm_common.cuh", "solid_dev.cuh", "p2p_dev.cuh", os.path.join("..", "..", "include", "rdc.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _metis() -> str:
    for root in ("/usr/local/cuda/targets/x86_64-linux/lib", "/usr/local/cuda-12.9/targets/x86_64-linux/lib"):
        p = os.path.join(root, "libmetis_static.a")
        if os.path.exists(p):
            return p
    raise RuntimeError("libmetis_static.a not found in the CUDA toolkit")


def up_to_date() -> bool:
    if not os.path.exists(SO):
        return False
    t = os.path.getmtime(SO)
    files = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(f) <= t for f in files if os.path.exists(f))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return SO
    if not all(os.path.exists(os.path.join(CSRC, f)) for f in SOURCES):
        if os.path.exists(SO):
            return SO
        raise RuntimeError("sources missing and no prebuilt librdcgpu.so")
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-ccbin", ccbin, "-Xcompiler", "-fPIC,-fopenmp,-O2", "-shared", "-o", SO]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, f) for f in SOURCES]
    cmd += [_metis(), "-lgomp", "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
