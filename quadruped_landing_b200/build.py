"""In-tree build of libqlnlp.so (the C-ABI CUDA library) with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the
resulting .so is git-ignored but travels to the GPU box with the working tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libqlnlp.so")
SOURCES = ["qlnlp.cu"]
HOST_SOURCES = ["hostrows.cpp"]        # plain C++ (g++): worker pool + row builder of the host-pointer path


def _deps():
    """Every file the library is built from: all of csrc/ plus the public header (a glob, so the list cannot drift)."""
    import glob
    out = []
    for pat in ("*.cu", "*.cuh", "*.h", "*.cpp"):
        out += glob.glob(os.path.join(CSRC, pat))
    out.append(os.path.join(_HERE, "..", "include", "qlnlp.h"))
    out.append(os.path.abspath(__file__))
    return out

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def _host_objects(prefix: str):
    """g++ -c of the plain C++ sources (the AVX-512 row writer sits behind a function-level target attribute, which
    nvcc's front end does not accept)."""
    objs = []
    for src in HOST_SOURCES:
        obj = f"{prefix}.{os.path.splitext(src)[0]}.o"
        res = subprocess.run(["g++", "-O3", "-std=c++17", "-fPIC", "-pthread", "-c", os.path.join(CSRC, src), "-o", obj],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
        objs.append(obj)
    return objs


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/qlnlp.cu -> libqlnlp.so if sources are newer than the library.

    Safe under concurrent callers (e.g. the ranks of a torchrun job): the build is serialised with a file lock
    and the library is replaced atomically."""
    import fcntl

    if not force and not is_stale():
        return LIB
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not is_stale():          # another process built it while we waited
            return LIB
        tmp = f"{LIB}.tmp.{os.getpid()}"
        objs = _host_objects(tmp)
        cmd = [nvcc_path(), *NVCC_FLAGS, "-Xcompiler", "-pthread", "-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES] + objs
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        for o in objs:
            os.unlink(o)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        os.replace(tmp, LIB)
        if verbose:
            print(res.stdout + res.stderr)
    return LIB


def build_variant(name: str, defines) -> str:
    """Compile an experimental variant (extra -D flags) to libqlnlp_<name>.so for A/B timing (tools/ab_bench.py)."""
    out = os.path.join(_HERE, f"libqlnlp_{name}.so")
    objs = _host_objects(out)
    cmd = [nvcc_path(), *NVCC_FLAGS, "-Xcompiler", "-pthread", *[f"-D{d}" for d in defines], "-o", out] + \
          [os.path.join(CSRC, s) for s in SOURCES] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    for o in objs:
        os.unlink(o)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
