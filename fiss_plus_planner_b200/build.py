"""Build libfissgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "fiss_abi.cu")
HDRS = [os.path.join(HERE, "csrc", "fiss_kernels.cuh"), os.path.join(ROOT, "include", "fiss_abi.h")]
OUT = os.path.join(HERE, "libfissgpu.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def build(force: bool = False, verbose: bool = False) -> str:
    deps = [SRC] + HDRS
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-o", OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libfissgpu.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
