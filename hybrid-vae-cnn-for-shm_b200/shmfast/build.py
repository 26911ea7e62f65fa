"""Build libshmfast.so in-tree with nvcc for sm_100a (B200).  `python -m shmfast.build [--force] [-v]`."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG.parent / "csrc"
INCLUDE = PKG.parent.parent / "include"
LIBDIR = PKG / "lib"
OBJDIR = CSRC / "build"
LIB = LIBDIR / "libshmfast.so"

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]
if os.environ.get("SHMFAST_PROF") == "1":          # role counters of the tensor-core scorers (csrc/vae_tc.cuh); part of the digest
    NVCC_FLAGS.append("-DSHM_TC_PROF")
NVCC_FLAGS += os.environ.get("SHMFAST_NVCC_EXTRA", "").split()          # experiment knobs (-D...), part of the digest


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [INCLUDE / "shmfast.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


STAMP = LIBDIR / "libshmfast.sha256"       # git-ignored: only ever written next to a freshly linked .so


def is_fresh() -> bool:
    return LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    OBJDIR.mkdir(exist_ok=True)
    stamp = STAMP
    dig = _digest()
    if not force and is_fresh():
        return LIB
    if stamp.exists():
        stamp.unlink()                           # a failed build must not leave a matching stamp behind
    cc = nvcc()

    def compile_one(src: Path):
        obj = OBJDIR / (src.stem + ".o")
        cmd = [cc, *NVCC_FLAGS, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        (OBJDIR / (src.stem + ".ptxas.log")).write_text(r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [cc, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(dig)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
