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


class FwdSeq(C.Structure):
    """oneprot_fwd_seq_t of include/oneprot_clip.h"""
    _fields_ = [("A", C.c_void_p), ("B_all", C.c_void_p), ("stats_rows", C.c_void_p), ("scale", C.c_void_p),
                ("stats", C.c_void_p), ("saved", C.c_void_p), ("zero_ptr", C.c_void_p), ("zero_bytes", C.c_size_t),
                ("sums", C.c_void_p), ("sums_mc", C.c_void_p), ("ag", C.POINTER(AgDesc)), ("ws", C.c_void_p),
                ("ws_bytes", C.c_size_t), ("stream", C.c_void_p),
                ("n", C.c_int), ("N", C.c_int), ("d", C.c_int), ("row_offset", C.c_int), ("mode", C.c_int),
                ("stats_rows_n", C.c_int), ("stats_off", C.c_int), ("E", C.c_void_p), ("lde", C.c_int)]


class BwdSeq(C.Structure):
    """oneprot_bwd_seq_t of include/oneprot_clip.h"""
    _fields_ = [("A", C.c_void_p), ("B_all", C.c_void_p), ("scale", C.c_void_p), ("stats", C.c_void_p),
                ("inv_rowsum", C.c_void_p), ("inv_colsum", C.c_void_p), ("g", C.c_void_p), ("dA", C.c_void_p),
                ("dB", C.c_void_p), ("ws", C.c_void_p), ("ws_bytes", C.c_size_t), ("panel_bytes", C.c_size_t),
                ("stream", C.c_void_p), ("side_stream", C.c_void_p), ("g_slot", C.c_void_p), ("g_slot_mc", C.c_void_p),
                ("dB_mc_mine", C.c_void_p), ("dB_out", C.c_void_p), ("seq", C.c_void_p),
                ("n", C.c_int), ("N", C.c_int), ("d", C.c_int), ("row_offset", C.c_int), ("mode", C.c_int),
                ("use_gsum", C.c_int), ("world", C.c_int), ("rank", C.c_int), ("want_a", C.c_int), ("want_b", C.c_int),
                ("g_on_side", C.c_int), ("E", C.c_void_p), ("lde", C.c_int)]


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
    "oneprot_clip_loss_finalize_ex": (_i, [_fp, _fp, _fp, _i, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _ip, _vp, _fp, _fp, _vp]),
    "oneprot_clip_rowcol_max": (_i, [_vp, _vp, _i, _i, _i, _fp, _fp, _fp, _vp, _sz, _vp]),
    "oneprot_augment_bf16": (_i, [_vp, _i, _i, _fp, _fp, _vp, _fp, _vp]),
    "oneprot_retrieval_ranks": (_i, [_vp, _vp, _i, _i, _fp, _fp, _fp, _vp, _sz, _vp]),
    "oneprot_clip_bwd_weights": (_i, [_fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _vp]),
    "oneprot_clip_dz_panel": (_i, [_vp, _vp, _i, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _vp, _i, _vp]),
    "oneprot_clip_fwd_sums_keep": (_i, [_vp, _vp, _i, _i, _i, _fp, _fp, C.POINTER(AgDesc), _fp, _fp, _vp, _sz, _vp, _i, _vp]),
    "oneprot_clip_dz_from_exp": (_i, [_vp, _i, _i, _i, _i, _fp, _fp, _fp, _vp]),
    "oneprot_gemm_bf16": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _fp, _fp, _vp, _i, _vp]),
    "oneprot_gemm_rowdot_scratch_bytes": (_sz, [_i, _i]),
    "oneprot_gemm_bf16_ex": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _fp, _fp, _vp, _i, _fp, _vp, _i, _fp, _vp]),
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
    "oneprot_siglip_fwd": (_i, [_vp, _vp, _i, _i, _i, _fp, _fp, _fp, _vp, _sz, _vp]),
    "oneprot_siglip_fwd_keep_scratch_bytes": (_sz, [_i, _i]),
    "oneprot_siglip_fwd_keep": (_i, [_vp, _vp, _i, _i, _i, _i, _fp, _fp, _fp, _fp, _vp, _sz, _vp, _i, _vp]),
    "oneprot_siglip_finalize": (_i, [_fp, _fp, _i, _fp, _fp, _fp, _vp]),
    "oneprot_siglip_dz_scratch_bytes": (_sz, [_i, _i]),
    "oneprot_siglip_dz_panel": (_i, [_vp, _vp, _i, _i, _i, _i, _fp, _fp, _fp, _fp, _vp, _i, _fp, _vp, _sz, _vp]),
    # projection-head row kernels (csrc/head_kernels.cu)
    "oneprot_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _fp, _fp, _i, _i, _i, _f, _vp]),
    "oneprot_layernorm_bwd_scratch_bytes": (_sz, [_i, _i]),
    "oneprot_layernorm_bwd": (_i, [_vp, _vp, _vp, _fp, _fp, _vp, _fp, _fp, _vp, _sz, _i, _i, _i, _vp]),
    "oneprot_gelu": (_i, [_vp, _vp, _vp, _sz, _i, _vp]),
    "oneprot_meanpool_fwd": (_i, [_vp, _fp, _vp, _fp, _i, _i, _i, _i, _i, _vp]),
    "oneprot_token_dot": (_i, [_vp, _vp, _i, _fp, _fp, _fp, _i, _i, _i, _i, _vp]),
    "oneprot_softmax_rows": (_i, [_fp, _fp, _i, _i, _vp]),
    "oneprot_softmax_rows_bwd": (_i, [_fp, _fp, _fp, _i, _i, _vp]),
    "oneprot_attnpool_bwd_x": (_i, [_vp, _fp, _fp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "oneprot_sum_slots_f32": (_i, [_fp, _i, _i, _i, _fp, _vp]),
    "oneprot_abs_mean_scratch_bytes": (_sz, [_sz]),
    "oneprot_abs_mean_fwd": (_i, [_vp, _sz, _sz, _i, _fp, _vp, _sz, _vp]),
    "oneprot_abs_mean_bwd": (_i, [_vp, _fp, _sz, _sz, _i, _vp, _vp]),
    "oneprot_meanpool_bwd": (_i, [_vp, _fp, _fp, _vp, _i, _i, _i, _i, _vp]),
    # host-side step sequencer + launch trace (csrc/clip_sequence.cu)
    "oneprot_seq_fwd_ws_bytes": (_sz, [_i, _i]),
    "oneprot_seq_fwd_begin": (_i, [C.POINTER(FwdSeq)]),
    "oneprot_seq_fwd_end": (_i, [C.POINTER(FwdSeq)]),
    "oneprot_seq_fwd": (_i, [C.POINTER(FwdSeq)]),
    "oneprot_seq_create": (_i, [C.POINTER(C.c_void_p)]),
    "oneprot_seq_destroy": (None, [_vp]),
    "oneprot_seq_bwd_ws_bytes": (_sz, [_i, _i, _i, _i, _i, _sz]),
    "oneprot_seq_bwd_ws_bytes_ex": (_sz, [_i, _i, _i, _i, _i, _sz, _i]),
    "oneprot_seq_bwd_panels": (_i, [_i, _i, _i, _sz, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "oneprot_seq_bwd_begin": (_i, [C.POINTER(BwdSeq)]),
    "oneprot_seq_bwd_main": (_i, [C.POINTER(BwdSeq)]),
    "oneprot_seq_bwd_end": (_i, [C.POINTER(BwdSeq)]),
    "oneprot_trace_begin": (None, [_i]),
    "oneprot_trace_end": (_sz, [C.c_char_p, _sz]),
    "oneprot_trace_note": (None, [C.c_char_p]),
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
