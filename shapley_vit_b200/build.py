"""Build ``libsvit.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m shapley_vit_b200.build [--force] [--verbose]

The shared library lands next to the sources (``shapley_vit_b200/csrc/libsvit.so``); it is
git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libsvit.so")
STAMP = os.path.join(CSRC, ".libsvit.stamp")
OBJDIR = os.path.join(CSRC, "build")

SOURCES = ["layout.cu", "aggregate.cu", "score.cu", "elementwise.cu", "attention.cu", "attention_mma.cu", "attention_tc.cu", "attention_tc_split.cu", "tma_util.cu",
           "gemm_simt.cu", "gemm_tc.cu", "forward.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libsvit cannot be built")


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _fingerprint() -> str:
    h = hashlib.sha256()
    names = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    names.append(os.path.join("..", "..", "include", "svit.h"))
    for n in names:
        with open(os.path.join(CSRC, n), "rb") as f:
            h.update(n.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == _fingerprint()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile and link; safe under several processes at once (one rank per GPU all importing the package): an
    exclusive file lock serialises the builders, the late comers find the library current, and the link goes to a
    temporary name that is renamed into place, so no process ever maps a half-written file."""
    if not force and is_current():
        return LIB
    import fcntl

    os.makedirs(OBJDIR, exist_ok=True)
    with open(os.path.join(CSRC, ".libsvit.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():      # another process built it while we waited
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
        if r.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    tmp = f"{LIB}.tmp.{os.getpid()}"
    cmd = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link of libsvit.so failed")
    os.replace(tmp, LIB)                         # atomic: readers see the old or the new file, never a partial one
    with open(STAMP + ".tmp", "w") as f:
        f.write(_fingerprint())
    os.replace(STAMP + ".tmp", STAMP)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
