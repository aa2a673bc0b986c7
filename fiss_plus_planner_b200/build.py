"""Build libfissgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "fiss_abi.cu")
import glob  # noqa: E402
# every header the translation unit can include: all of csrc/*.cuh + the public ABI header
HDRS = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh"))) + sorted(glob.glob(os.path.join(ROOT, "include", "*.h")))
OUT = os.path.join(HERE, "libfissgpu.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def build(force: bool = False, verbose: bool = False, out: str = OUT, extra_flags=()) -> str:
    """``out`` / ``extra_flags`` build a variant beside the product library (e.g. -DFISS_GRID_MIN_CTAS=2 for an
    A/B run; load it with FISSGPU_LIB=<path>)."""
    deps = [SRC] + HDRS
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-I", os.path.join(ROOT, "include"), "-o", out, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libfissgpu.so")
    return out


if __name__ == "__main__":
    flags = [a for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv or bool(outs), verbose="--quiet" not in sys.argv,
                out=os.path.abspath(outs[0]) if outs else OUT, extra_flags=flags))
