"""Build libsmalfit.so (CUDA, sm_100a) in-tree with nvcc.

    python -m smalify_b200.build

The library travels to the GPU box with the repo snapshot (``*.so`` is git-ignored
but not gpurun-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsmalfit.so")
SOURCES = ["smalfit_kernels.cu", "smalfit_capi.cu"]
PUBLIC_HEADER = os.path.join(HERE, "..", "include", "smalfit.h")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsmalfit.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    # every file of csrc/ (sources and the headers they include) + the public header
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + [PUBLIC_HEADER, os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def source_hash() -> str:
    """sha256 (16 hex digits) over the CUDA sources, their headers and the public header: identifies the code a
    measurement was taken on (nvcc's output is not bit-reproducible, the library file's own hash is not usable)."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    for path in [os.path.join(CSRC, f) for f in files] + [PUBLIC_HEADER]:
        with open(path, "rb") as fh:
            h.update(os.path.basename(path).encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """Build the library.  `defines` / `out` produce an experiment variant beside the product
    library (tools/ab_bench.py); the product is always the default build."""
    if out is None and not defines and not force and not is_stale():
        return LIB
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
           "-diag-suppress", "1886"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-D" + d for d in defines]
    target = out or LIB
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", target]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
