"""Build libast_b200.so in-tree with nvcc for sm_100a only.

The shared library is a plain C-ABI object (include/ast_b200.h): no torch headers, no pybind.
It is written next to this file so that it travels to the GPU box with the repo snapshot
(``*.so`` is git-ignored, not gpurun-ignored).  ``python -m arbitrarystyletransfer_b200._build``
rebuilds it; ``__graft_entry__.build()`` calls :func:`build`.
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
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
BUILD_DIR = os.path.join(HERE, "build")
LIB_PATH = os.path.join(HERE, "libast_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: libast_b200.so cannot be built (no CPU fallback exists)")


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path: str, extra: str) -> str:
    h = hashlib.sha256(extra.encode())
    for dep in [path] + sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))
    ) + [os.path.join(INCLUDE, "ast_b200.h")]:
        with open(dep, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ and link libast_b200.so.  Returns the library path.
    AST_KERNEL_DEBUG=1 in the environment compiles the kernels' wait counters and bottleneck-elimination flags in
    (AST_CONV_DEBUG / AST_CONV_DBGFLAGS / AST_PW_DBGFLAGS then work; the kernels are 15-40 % slower)."""
    nvcc = _nvcc()
    os.makedirs(BUILD_DIR, exist_ok=True)
    flags = list(NVCC_FLAGS)
    if os.environ.get("AST_KERNEL_DEBUG", "0") not in ("", "0"):
        flags += ["-DAST_KERNEL_DEBUG=1"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    objs, jobs = [], []
    for src in sources():
        obj = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + ".o")
        stamp = obj + ".sha"
        dig = _digest(src, " ".join(flags))
        objs.append(obj)
        fresh = (not force and os.path.isfile(obj) and os.path.isfile(stamp)
                 and open(stamp).read() == dig)
        if not fresh:
            jobs.append((src, obj, stamp, dig))

    def compile_one(job):
        src, obj, stamp, dig = job
        cmd = [nvcc] + flags + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        with open(stamp, "w") as f:
            f.write(dig)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or force or not os.path.isfile(LIB_PATH):
        cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-cudart", "static",
               "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
