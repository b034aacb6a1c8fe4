"""Build libbposd_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbposd_b200.so")
SOURCES = ["bposd_capi.cu"]
DEPS = ["bposd_capi.cu", "bposd_kernels.cuh", "bp_fast_kernel.cuh", "bp_cluster_kernel.cuh", "osd_panel_kernel.cuh", os.path.join("..", "..", "include", "bposd_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # fp64 mode must not contract a*b+c: bit-exactness with the reference order
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None, fast: bool = False) -> str:
    """defines/out: experimental A/B builds (e.g. defines=("BPOSD_OLDMIN",), out="/path/lib_b.so").
    fast: -split-compile 0 (2.5x quicker; register allocation differs slightly, so release builds do not use it)."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + (["-split-compile", "0"] if fast else []) + \
          [f"-D{d}" for d in defines] + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out or LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libbposd_b200.so")
    return out or LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
