"""Builds semstereo_b200/libsemstereo_b200.so from csrc/*.cu with nvcc for sm_100a (in-tree, no JIT cache).

    python -m semstereo_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the tree.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libsemstereo_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode())
            h.update(open(os.path.join(CSRC, f), "rb").read())
    inc = os.path.join(os.path.dirname(HERE), "include", "semstereo_b200.h")
    if os.path.exists(inc):
        h.update(open(inc, "rb").read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Idempotent and safe under torchrun: the digest check, compile and link run under an inter-process file lock (every rank
    of a fresh tree would otherwise write the same objects), and the library and its stamp are moved into place atomically, so no
    process can dlopen a half-written .so."""
    import fcntl
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "digest.txt")
    dig = _digest()

    def fresh():
        return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig

    if not force and fresh():
        return LIB
    with open(os.path.join(OBJ, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and fresh():          # another process built it while we waited
                return LIB
            return _build_locked(dig, stamp, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(dig: str, stamp: str, verbose: bool) -> str:
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}; cannot build {LIB}")

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC, *FLAGS, "-I", CSRC, "-I", os.path.join(os.path.dirname(HERE), "include"), "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr.strip():
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    tmp = f"{LIB}.{os.getpid()}.tmp"
    cmd = [NVCC, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if os.path.exists(stamp):
        os.remove(stamp)                       # never a fresh stamp next to an old library
    os.replace(tmp, LIB)
    with open(stamp + ".tmp", "w") as f:
        f.write(dig)
    os.replace(stamp + ".tmp", stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
