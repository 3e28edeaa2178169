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
HEADERS = ["layout.h", "qlnlp_kernels.cuh", "rk4_dual_gen.h", os.path.join("..", "..", "include", "qlnlp.h")]

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
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


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
        cmd = [nvcc_path(), *NVCC_FLAGS, "-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        os.replace(tmp, LIB)
        if verbose:
            print(res.stdout + res.stderr)
    return LIB


def build_variant(name: str, defines) -> str:
    """Compile an experimental variant (extra -D flags) to libqlnlp_<name>.so for A/B timing (tools/ab_bench.py)."""
    out = os.path.join(_HERE, f"libqlnlp_{name}.so")
    cmd = [nvcc_path(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
