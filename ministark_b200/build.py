"""Build libministark.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels to the GPU box)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libministark.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-shared",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    hdr = os.path.join(os.path.dirname(HERE), "include", "ministark.h")
    return any(os.path.getmtime(s) > t for s in sources() + [hdr])


def build(force: bool = False, verbose: bool = False, out: str | None = None, defines: list[str] | None = None) -> str:
    """out/defines: experimental A/B builds (MINISTARK_LIB=<out> selects one at run time)"""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = ([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [f"-D{d}" for d in (defines or [])]
           + ["-o", out or LIB, os.path.join(CSRC, "ministark.cu")])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libministark.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return out or LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
