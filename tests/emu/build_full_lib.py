"""Builds the WHOLE library for the CPU: clip_kernels.cu + head_kernels.cu + clip_sequence.cu, kernels and
C ABI host functions alike (launches rewritten by translate.py, CUDA runtime stubbed by host_emu.h).  The
result exports the symbols of include/oneprot_clip.h, so `ONEPROT_LIB=<it>` makes oneprot_b200._lib bind the
emulated library: Python host + ctypes + C host code + kernel source then all run without a GPU.
Test infrastructure only.

    python tests/emu/build_full_lib.py [out.so]
"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "oneprot_b200", "csrc")
sys.path.insert(0, HERE)
import translate  # noqa: E402


def build(out=None):
    stamp = str(int(max(os.path.getmtime(os.path.join(d, f)) for d in (HERE, CSRC, os.path.join(ROOT, "include")) for f in os.listdir(d))))
    cache = os.path.join(tempfile.gettempdir(), "oneprot_emu_cache_" + stamp)
    os.makedirs(cache, exist_ok=True)
    out = out or os.path.join(cache, "liboneprot_clip_emu.so")
    if os.path.exists(out):
        return out
    gen = os.path.join(cache, f"gen_{os.getpid()}")
    os.makedirs(gen, exist_ok=True)
    tus = []
    for f in ("clip_kernels.cu", "head_kernels.cu", "clip_sequence.cu"):       # one translation unit each, as in the CUDA build
        text = translate.translate(open(os.path.join(CSRC, f)).read())
        text = text.replace('"../../include/oneprot_clip.h"', '"' + os.path.join(ROOT, "include", "oneprot_clip.h") + '"')
        tu = os.path.join(gen, f.replace(".cu", "_emu.cpp"))
        open(tu, "w").write('#define ONEPROT_KERNEL_EMULATION 1\n#define ONEPROT_HOST_EMULATION 1\n#include "host_emu.h"\n' + text)
        tus.append(tu)
    tmp = out + f".{os.getpid()}.tmp"
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-w", "-I/usr/local/cuda/include", "-I" + HERE, "-I" + CSRC,
           "-o", tmp, *tus]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(p.stderr[-6000:])
    os.replace(tmp, out)
    return out


if __name__ == "__main__":
    print(build(sys.argv[1] if len(sys.argv) > 1 else None))
