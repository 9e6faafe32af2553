"""In-tree build of libfdgraph.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfdgraph.so")
SOURCES = ["fdg_capi.cu", "fdg_lower.cpp", "fdg_jit.cpp", "fdg_lgjit.cpp"]
HEADERS = ["fdg_vm.cuh", "fdg_isa.h", "fdg_lower.h", "fdg_jit.h", os.path.join("..", "..", "include", "fdgraph.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",  # belt and braces: the kernels use __dmul_rn/__dadd_rn, which are never contracted
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libfdgraph.so cannot be built (there is no CPU fallback)")
    return exe


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into feynmandiagram.jl_b200/libfdgraph.so; returns the path."""
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES + ["-lnvptxcompiler_static", "-lnvJitLink_static", "-ldl", "-lpthread"]
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stdout + proc.stderr)
    return LIB
