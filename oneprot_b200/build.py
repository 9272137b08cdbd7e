"""Builds the CUDA extension in-tree with nvcc for sm_100a (no torch cpp_extension, no JIT cache).

    python -m oneprot_b200.build          # -> oneprot_b200/liboneprot_clip.so
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRCS = [os.path.join(HERE, "csrc", "clip_kernels.cu"),     # kernels + one entry point per kernel
        os.path.join(HERE, "csrc", "clip_sequence.cu"),    # host-side step sequencer + launch trace
        os.path.join(HERE, "csrc", "head_kernels.cu")]     # projection-head row kernels (pooling, LayerNorm, GELU)
DEPS = SRCS + [os.path.join(HERE, "csrc", "ptx.cuh"), os.path.join(HERE, "csrc", "vector_kernels.cuh"), os.path.join(HERE, "csrc", "host_trace.h"),
               os.path.join(HERE, "..", "include", "oneprot_clip.h")]
LIB = os.path.join(HERE, "liboneprot_clip.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build the oneprot_b200 CUDA extension")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    cmd = [find_nvcc(), *NVCC_FLAGS, "-o", LIB, *SRCS]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
