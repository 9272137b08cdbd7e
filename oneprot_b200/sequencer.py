"""Python side of the host-side step sequencer (``csrc/clip_sequence.cu``).

One ``ClipLoss`` fwd+bwd through ``clip_loss.py`` is ~35 host operations; here a whole phase of the
step is ONE call into ``liboneprot_clip.so`` working out of one workspace tensor: 1 call for the
forward and 1 for the backward on a single GPU, 2 + 3 around the symmetric-memory barriers with the
NVLS exchange provider.  The C side issues the same launches in the same order on the same streams
as ``_ClipLossFunction._forward_impl/_backward_impl`` (``tests/test_sequencer_cpu.py`` compares the
two launch traces), so the numerics are those of the Python path.

Default since round 2 (``ClipLoss(host_sequencer=True)``; validated on B200 at 1, 2 and 8 GPUs: bit-identical to
the Python host).  Scope (everything else stays on the Python path): bf16 features with
d % 8 == 0, ``logit_scale`` without gradient, one pass over both softmax directions (world 1;
``local_loss=False``; ``local_loss=True`` with ``gather_with_grad=True``), and for world > 1 the NVLS
provider with the all-gather fused into the forward kernel and a side stream for the exchanges.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import check

HDR = 16            # ONEPROT_SAVED_HEADER_FLOATS
STATS_AT = 4        # ONEPROT_SAVED_STATS_AT
FLAG_AT = 8         # ONEPROT_SAVED_FLAG_AT


def enabled(cfg) -> bool:
    return bool(cfg.get("host_sequencer"))


def eligible(cfg, ops, comm, scale_requires_grad: bool) -> bool:
    if ops.split or ops.pad or scale_requires_grad or cfg.get("robust", "off") != "off":
        return False
    W = cfg["world_size"]
    if W == 1:
        return comm.name == "local"
    if comm.name != "nvls" or (cfg["local_loss"] and not cfg["gather_with_grad"]):
        return False
    return comm.seq_ready(ops.n, ops.A.device)


class _State:
    """What the backward needs from the forward (kept on ctx)."""
    __slots__ = ("saved", "B_all_ptr", "B_keep", "token", "mode", "scale_dev", "E", "hold")


def forward(ctx, ops, scale_dev, cfg, comm, K, keep=None):
    """-> (loss32 view, flag view); fills ctx.seq.  keep: bf16 panel that receives the exponentials
    (stored-exponentials backward, clip_loss.py) or None."""
    lib = _lib.load()
    K._need_cuda(ops.A, ops.B)        # no CPU fallback: fail like every kernel wrapper does
    W, rank = cfg["world_size"], cfg["rank"]
    n, d = ops.n, ops.d
    N, off = W * n, rank * n
    dev = ops.A.device
    mode = K.MODE_LOCAL if (W > 1 and cfg["local_loss"]) else K.MODE_GLOBAL
    saved = torch.empty(HDR + 2 * N, dtype=torch.float32, device=dev)
    ws_bytes = int(lib.oneprot_seq_fwd_ws_bytes(n, N))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    sp = saved.data_ptr()
    f = _lib.FwdSeq()
    f.A, f.scale, f.saved = ops.A.data_ptr(), scale_dev.data_ptr(), sp
    f.ws, f.ws_bytes, f.stream = ws.data_ptr(), ws_bytes, K.current_stream_handle()
    f.n, f.N, f.d, f.row_offset, f.mode = n, N, d, off, mode
    st = _State()
    st.saved, st.mode, st.scale_dev, st.E = saved, mode, scale_dev, keep
    if keep is not None:
        f.E, f.lde = keep.data_ptr(), keep.stride(0)
    if W == 1:
        f.B_all = f.stats_rows = ops.B.data_ptr()
        f.stats_rows_n, f.stats_off = N, 0
        f.stats = sp + 4 * STATS_AT
        check(lib.oneprot_seq_fwd(C.byref(f)), "oneprot_seq_fwd")
        st.B_all_ptr, st.B_keep, st.token, st.hold = ops.B.data_ptr(), ops.B, None, None
    else:
        x = comm.seq_forward_desc(ops, rank, W, sp + 4 * STATS_AT)
        ag = _lib.AgDesc(*x["ag"])
        f.B_all, f.stats_rows = x["B_all"], ops.B.data_ptr()
        f.stats_rows_n, f.stats_off = n, 0
        f.stats = x["stats"]
        f.zero_ptr, f.zero_bytes = x["zero_ptr"], x["zero_bytes"]
        f.sums, f.sums_mc = x["sums"], x["sums_mc"]
        f.ag = C.pointer(ag)
        check(lib.oneprot_seq_fwd_begin(C.byref(f)), "oneprot_seq_fwd_begin")
        comm._barrier()                       # every rank's partial sums are in place
        check(lib.oneprot_seq_fwd_end(C.byref(f)), "oneprot_seq_fwd_end")
        st.B_all_ptr, st.B_keep, st.token = x["B_all"], None, x["token"]
        st.hold = comm.hold_for_backward(x["B_view"]) if any(ctx.needs_input_grad[:2]) else None
    ctx.seq = st
    return saved[0:1], saved[FLAG_AT:FLAG_AT + 1].view(torch.int32)


def backward(ctx, g_loss, cfg, ops, comm, K):
    """-> (grad_a, grad_b) in bf16 (None where not needed)."""
    lib = _lib.load()
    st = ctx.seq
    W, rank = cfg["world_size"], cfg["rank"]
    n, d = ops.n, ops.d
    N, off = W * n, rank * n
    dev = ops.A.device
    need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    g32 = (torch.zeros(1, dtype=torch.float32, device=dev) if g_loss is None
           else g_loss.detach().to(device=dev, dtype=torch.float32).reshape(1))
    want_a, want_b = bool(need_a), bool(need_b or W > 1)    # with W > 1 every rank enters the exchange
    E, st.E = st.E, None                 # consumed (overwritten) by this backward; a second one recomputes
    ws_bytes = int(lib.oneprot_seq_bwd_ws_bytes_ex(n, N, d, W, int(want_b), cfg["panel_bytes"], int(E is not None)))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    dA = torch.empty(n, d, dtype=torch.bfloat16, device=dev) if want_a else None
    sp = st.saved.data_ptr()
    q = _lib.BwdSeq()
    q.A, q.scale, q.stats = ops.A.data_ptr(), st.scale_dev.data_ptr(), sp + 4 * STATS_AT
    q.inv_rowsum, q.inv_colsum, q.g = sp + 4 * HDR, sp + 4 * (HDR + N), g32.data_ptr()
    q.dA = dA.data_ptr() if want_a else None
    q.ws, q.ws_bytes, q.panel_bytes = ws.data_ptr(), ws_bytes, cfg["panel_bytes"]
    q.stream = K.current_stream_handle()
    q.n, q.N, q.d, q.row_offset, q.mode = n, N, d, off, st.mode
    q.use_gsum, q.world, q.rank = int(cfg["gather_with_grad"]), W, rank
    q.want_a, q.want_b = int(want_a), int(want_b)
    if E is not None:
        q.E, q.lde = E.data_ptr(), E.stride(0)
    if W == 1:
        dB = torch.empty(N, d, dtype=torch.bfloat16, device=dev) if want_b else None
        q.B_all = st.B_all_ptr
        q.dB = dB.data_ptr() if want_b else None
        check(lib.oneprot_seq_bwd_main(C.byref(q)), "oneprot_seq_bwd_main")
        return dA, dB
    B_keep = comm.b_all_for_backward(ops, None, st.token, rank, W, hold=st.hold)   # None: the gathered operand is still in place
    q.B_all = B_keep.data_ptr() if B_keep is not None else st.B_all_ptr
    side = comm.side_stream(dev)
    dB_out = torch.empty(n, d, dtype=torch.bfloat16, device=dev)
    x = comm.seq_backward_desc(n, d, rank, st.token)
    q.dB, q.dB_mc_mine, q.dB_out = x["dB"], x["dB_mc_mine"], dB_out.data_ptr()
    q.g_slot, q.g_slot_mc = x["g_slot"], x["g_slot_mc"]
    q.side_stream, q.seq = side.cuda_stream, comm.seq_handle()
    q.g_on_side = int(st.mode == K.MODE_GLOBAL)
    check(lib.oneprot_seq_bwd_begin(C.byref(q)), "oneprot_seq_bwd_begin")
    if q.g_on_side:
        with torch.cuda.stream(side):
            comm._barrier()                   # every rank's one-hot gradient is in its slot
    else:
        comm._barrier()
    check(lib.oneprot_seq_bwd_main(C.byref(q)), "oneprot_seq_bwd_main")
    with torch.cuda.stream(side):
        comm._barrier()                       # every rank's partial dB is written
    check(lib.oneprot_seq_bwd_end(C.byref(q)), "oneprot_seq_bwd_end")
    return dA, (dB_out if need_b else None)
