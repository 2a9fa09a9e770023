"""Build libp3p.so (the C-ABI CUDA library of this package) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
# P3P_LIB selects another build of the same library (kernel experiments: tools/build_variant.py); the product default is in-tree
LIB_PATH = os.environ.get("P3P_LIB") or os.path.join(_HERE, "libp3p.so")
SOURCES = ["capi.cu", "voxelize.cu", "pfn.cu", "patch_embed.cu", "las.cu", "conv3x3.cu", "pfn_train.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--cudart", "static",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(exe):
        raise RuntimeError("nvcc not found: libp3p.so cannot be built")
    return exe


def is_stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "p3p.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, out: str = None, extra_flags=None) -> str:
    out = out or LIB_PATH
    if out == LIB_PATH and not force and not is_stale():
        return LIB_PATH
    extra = list(extra_flags) if extra_flags is not None else os.environ.get("P3P_EXTRA_NVCC_FLAGS", "").split()
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-I", INCLUDE, "-I", CSRC, "-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
