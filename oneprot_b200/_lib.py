"""ctypes binding of the C ABI declared in include/oneprot_clip.h.

There is deliberately no fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# ONEPROT_LIB overrides the library path (A/B timing of kernel variants during development)
LIB_PATH = os.environ.get("ONEPROT_LIB") or os.path.join(HERE, "liboneprot_clip.so")

_vp, _fp, _ip = C.c_void_p, C.c_void_p, C.c_void_p   # all raw device pointers travel as void*
_i, _f, _sz = C.c_int, C.c_float, C.c_size_t

class AgDesc(C.Structure):
    """oneprot_ag_t of include/oneprot_clip.h"""
    _fields_ = [("src", C.c_void_p), ("dst_mc", C.c_void_p), ("counters", C.c_void_p), ("flags_mc", C.c_void_p),
                ("flags", C.c_void_p), ("stats_mc", C.c_void_p), ("stats_all", C.c_void_p), ("stats_out", C.c_void_p),
                ("epoch", C.c_uint), ("rank", C.c_int), ("world", C.c_int), ("chunks", C.c_int),
                ("rows_per_rank", C.c_int)]


# name -> (restype, argtypes); must list every symbol include/oneprot_clip.h declares
SIGNATURES = {
    "oneprot_abi_version": (_i, []),
    "oneprot_last_error": (C.c_char_p, []),
    "oneprot_device_check": (_i, [_i]),
    "oneprot_num_sms": (_i, []),
    "oneprot_launch_count": (C.c_longlong, []),
    "oneprot_launch_count_reset": (None, []),
    "oneprot_clip_rowstats": (_i, [_vp, _vp, _i, _i, _i, _i, _fp, _fp, _vp]),
    "oneprot_clip_fwd_scratch_bytes": (_sz, [_i, _i]),
    "oneprot_clip_fwd_sums": (_i, [_vp, _vp, _i, _i, _i, _fp, _fp, _fp, _fp, _vp, _sz, _vp]),
    "oneprot_clip_fwd_sums_ag": (_i, [_vp, _vp, _i, _i, _i, _fp, _fp, C.POINTER(AgDesc), _fp, _fp, _vp, _sz, _vp]),
    "oneprot_clip_loss_finalize": (_i, [_fp, _fp, _fp, _i, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _ip, _vp, _vp]),
    "oneprot_clip_bwd_weights": (_i, [_fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _vp]),
    "oneprot_clip_dz_panel": (_i, [_vp, _vp, _i, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _vp, _i, _vp]),
    "oneprot_gemm_bf16": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _fp, _fp, _vp, _i, _vp]),
    "oneprot_gemm_rowdot_scratch_bytes": (_sz, [_i, _i]),
    "oneprot_gemm_bf16_ex": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _fp, _fp, _vp, _i, _fp, _vp, _i, _fp, _vp]),
    "oneprot_gemm_bf16_push": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _fp, _i, _fp, C.POINTER(C.c_void_p), _i, _i, _i, _i, _vp]),
    "oneprot_sum_slots_bf16": (_i, [_vp, _i, _sz, _vp, _vp]),
    "oneprot_rowdot_bf16": (_i, [_vp, _i, _vp, _i, _i, _i, _fp, _vp]),
    "oneprot_sum_f32": (_i, [_fp, _i, _fp, _vp]),
    "oneprot_l2norm_scale_fwd": (_i, [_vp, _vp, _fp, _i, _i, _i, _fp, _f, _vp]),
    "oneprot_l2norm_scale_bwd": (_i, [_vp, _vp, _fp, _vp, _fp, _i, _i, _i, _fp, _f, _vp]),
    "oneprot_scale_rows": (_i, [_vp, _vp, _i, _i, _i, _fp, _vp]),
    "oneprot_rowdot": (_i, [_vp, _vp, _i, _i, _i, _fp, _vp]),
    "oneprot_mc_store": (_i, [_vp, _vp, _sz, _vp]),
    "oneprot_mc_allreduce_f32": (_i, [_fp, _fp, _i, _i, _vp]),
    "oneprot_mc_reduce_bf16": (_i, [_vp, _vp, _sz, _vp]),
    "oneprot_split_fp32": (_i, [_fp, _vp, _i, _i, _i, _i, _vp]),
}

_lib = None


class OneProtKernelError(RuntimeError):
    pass


def load():
    """Load liboneprot_clip.so (built by oneprot_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OneProtKernelError(
            f"{LIB_PATH} is missing: build it with `python -m oneprot_b200.build` "
            "(there is no CPU or PyTorch fallback for the ClipLoss path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.oneprot_abi_version() != 1:
        raise OneProtKernelError("liboneprot_clip.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().oneprot_last_error().decode("utf-8", "replace")
        raise OneProtKernelError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Raw device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
