"""Build libbposd_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

The library is several translation units compiled in parallel and cached by content hash under
``bp_osd_b200/csrc/build/`` (git-ignored): the host side + OSD / harness kernels (``bposd_capi.cu``), and one unit
per (precision, degree class) of the two specialised BP kernels (``bp_fast_inst.cu``, ``bp_cluster_inst.cu``).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libbposd_b200.so")
_INC = os.path.join("..", "..", "include")
HEADERS = ["bposd_kernels.cuh", "bp_fast_kernel.cuh", "bp_cluster_kernel.cuh", "osd_reg_kernel.cuh", "osd_cluster_kernel.cuh",
           os.path.join(_INC, "bposd_math.h"), os.path.join(_INC, "bposd_b200.h")]
# headers each source actually includes (the object cache is keyed by their contents)
DEPS = {"bposd_capi.cu": HEADERS,
        "bp_fast_inst.cu": ["bposd_kernels.cuh", "bp_fast_kernel.cuh", os.path.join(_INC, "bposd_math.h")],
        "bp_cluster_inst.cu": ["bposd_kernels.cuh", "bp_fast_kernel.cuh", "bp_cluster_kernel.cuh", os.path.join(_INC, "bposd_math.h")]}
CLASSES = [(4, 2), (6, 3), (8, 4), (16, 8)]  # (max row degree, max column degree) classes of the specialised BP kernels

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # fp64 mode must not contract a*b+c: bit-exactness with the reference order
    "-Xcompiler", "-fPIC",
]


def units(only_class=None):
    """(object name, source, extra defines) of every translation unit."""
    out = [("capi", "bposd_capi.cu", [])]
    for real, tag in (("double", "f64"), ("float", "f32")):
        for dc, dv in CLASSES:
            if only_class and (dc, dv) != only_class:
                continue
            d = [f"-DBPOSD_INST_REAL={real}", f"-DBPOSD_INST_DC={dc}", f"-DBPOSD_INST_DV={dv}"]
            out.append((f"fast_{tag}_{dc}", "bp_fast_inst.cu", d))
            out.append((f"cluster_{tag}_{dc}", "bp_cluster_inst.cu", d))
        if not only_class or only_class == (7, 4):  # rows of 7 slots: a class of the cluster kernel alone
            out.append((f"cluster_{tag}_7", "bp_cluster_inst.cu", [f"-DBPOSD_INST_REAL={real}", "-DBPOSD_INST_DC=7", "-DBPOSD_INST_DV=4"]))
    return out


def _digest(src, flags):
    h = hashlib.sha256()
    h.update(" ".join(flags).encode())
    for f in [src] + DEPS[src]:
        p = os.path.join(CSRC, f)
        if os.path.exists(p):
            with open(p, "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()[:16]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = ["bposd_capi.cu", "bp_fast_inst.cu", "bp_cluster_inst.cu"] + HEADERS
    return any(os.path.exists(os.path.join(CSRC, d)) and os.path.getmtime(os.path.join(CSRC, d)) > t for d in srcs)


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None, fast: bool = False,
          jobs: int | None = None) -> str:
    """defines/out: experimental A/B builds (e.g. defines=("BPOSD_OLDMIN",), out="/path/lib_b.so").
    fast: -split-compile 0 (quicker; register allocation differs slightly, so release builds do not use it).
    force: relink even if the library is newer than the sources (objects are still taken from the hash cache)."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    common = NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + (["-split-compile", "0"] if fast else []) + \
        [f"-D{d}" for d in defines]

    def compile_one(u):
        name, src, extra = u
        flags = common + extra
        obj = os.path.join(OBJ, f"{name}.{_digest(src, flags)}.o")
        if os.path.exists(obj) and not verbose:
            return obj, ""
        for old in os.listdir(OBJ):  # one cached object per unit name (and per A/B define set)
            if old.startswith(name + ".") and old.endswith(".o") and not defines:
                os.remove(os.path.join(OBJ, old))
        res = subprocess.run([nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj + ".tmp"],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src} ({name}):\n{res.stdout}{res.stderr}")
        os.replace(obj + ".tmp", obj)
        return obj, res.stdout + res.stderr

    only = None
    if os.environ.get("BPOSD_DEV_CLASS"):  # tuning builds: "6,3" compiles that degree class alone
        a, b = os.environ["BPOSD_DEV_CLASS"].split(",")
        only = (int(a), int(b))
        common = common + ["-DBPOSD_DEV_CLASS_ONLY=1", f"-DBPOSD_DEV_DC={only[0]}"]
    with ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as pool:
        results = list(pool.map(compile_one, units(only)))
    if verbose:
        sys.stderr.write("".join(log for _, log in results))
    objs = [o for o, _ in results]
    target = out or LIB
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", "-fPIC"]
                         + objs + ["-o", target + ".tmp"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed linking libbposd_b200.so:\n" + res.stdout + res.stderr)
    os.replace(target + ".tmp", target)
    return target


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
