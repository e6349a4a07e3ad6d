"""Build libcrispy_ns.so (the sm_100a kernels + C ABI) in-tree with nvcc."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcrispy_ns.so")
SOURCES = ["crispy_ns.cu", "ns_host.cpp"]
HEADERS = ["ns_common.h", "ns_simt.h", "ns_pipe.cuh", "ns_pitch7.cuh", "ns_rnn_tc5.cuh", "ns_host.h", "crispy_ns_abi.inc", os.path.join("..", "..", "include", "crispy_ns.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set $NVCC)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> crispy_b200/libcrispy_ns.so"""
    if not force and not is_stale():
        return LIB
    cmd = [
        nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        # exactness contract (ns_pipe.cuh): no implicit a*b+c contraction; fmaf() is spelled out where allowed
        "-fmad=false",
        "-Xcompiler", "-fPIC", "-shared", "-o", LIB,
    ] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
