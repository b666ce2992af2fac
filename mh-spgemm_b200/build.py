"""Build the CUDA library in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python mh-spgemm_b200/build.py            # -> mh-spgemm_b200/libmhb_spgemm.so

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmhb_spgemm.so")
SOURCES = ["mhb_capi.cu", "mhb_compat.cu", "mhb_shard.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
    "-I", os.path.join(HERE, "..", "include"),
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(HERE, "..", "include", f) for f in os.listdir(os.path.join(HERE, "..", "include"))]
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = ["nvcc", *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", LIB, *srcs, "-ldl"]
    subprocess.check_call(cmd)
    return LIB


def build_compat_driver() -> str:
    """The reference-style C++ driver of tests/cpp (links the shim in mhb_compat.cu)."""
    root = os.path.join(HERE, "..")
    src = os.path.join(root, "tests", "cpp", "compat_driver.cu")
    out = os.path.join(root, "tests", "cpp", "compat_driver")
    if os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(src), os.path.getmtime(LIB)):
        return out
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", os.path.join(root, "include"),
                           "-o", out, src, "-L", HERE, "-lmhb_spgemm", "-Xlinker", "-rpath", "-Xlinker",
                           "$ORIGIN/../../mh-spgemm_b200"])
    return out


def build_shard_driver() -> str:
    """tests/cpp/shard_driver.cpp: a plain C++ (g++, no CUDA headers) two-process caller of mhb_shard_*."""
    root = os.path.join(HERE, "..")
    src = os.path.join(root, "tests", "cpp", "shard_driver.cpp")
    out = os.path.join(root, "tests", "cpp", "shard_driver")
    if os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(src), os.path.getmtime(LIB)):
        return out
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(root, "include"), "-o", out, src,
                           "-L", HERE, "-lmhb_spgemm", "-Wl,-rpath,$ORIGIN/../../mh-spgemm_b200"])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
