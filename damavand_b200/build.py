"""In-tree build of libdamavand_b200.so with nvcc for sm_100a.

Replaces /root/reference/build.rs + damavand-gpu/CMakeLists.txt (cmake, no -arch flag at all).
The Rust-side equivalent (ffi/build.rs) runs the same nvcc command line.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdamavand_b200.so")
SOURCES = ["kernels.cu", "engine.cu", "compat.cu", "planner.cpp"]
HEADERS = ["kernels.h", "tile_core.cuh", "planner.h", "nccl_dyn.h",
           os.path.join("..", "..", "include", "damavand_b200.h"),
           os.path.join("..", "..", "include", "damavand_gpu_compat.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3", "-shared", "-cudart", "static",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; cannot build libdamavand_b200.so")
    return p


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    # development: DVD_BUILD_DEFS="-DX -DY" DVD_BUILD_OUT=path builds an A/B variant next to the product library
    out = os.environ.get("DVD_BUILD_OUT") or LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + os.environ.get("DVD_BUILD_DEFS", "").split() + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libdamavand_b200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
