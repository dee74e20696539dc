"""Build libusim.so (sm_100a only) in-tree with nvcc.  No JIT, no fallbacks."""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libusim.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + [
        os.path.join(_HERE, "..", "include", "usim.h")
    ]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False, out: str = LIB, extra=()) -> str:
    if not force and out == LIB and not needs_build():
        return LIB
    cmd = [
        NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math",
        "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v" if verbose else "-O3",
        "-o", out, os.path.join(CSRC, "usim.cu"), *extra,
    ]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libusim.so")
    if verbose:
        print(res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
