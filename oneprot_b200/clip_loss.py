"""Drop-in ``ClipLoss`` / ``gather_features`` for OneProt on B200.

Mirrors the reference interface (klemens-floege/oneprot ``src/models/components/loss.py``):
``ClipLoss(local_loss, gather_with_grad, cache_labels, rank, world_size, use_horovod)`` and
``forward(modality_features, sequence_features, logit_scale=1.0, output_dict=False)``
(loss.py:49-114), plus ``gather_features`` (loss.py:19-46).  The module registers no parameters or
buffers and touches no process group in ``__init__`` (the reference builds it before DDP
initialises, ``src/models/oneprot_module.py:48-56`` / ``src/train.py:44,88``).

What runs where
  * every floating-point operation on the path is a kernel of ``liboneprot_clip.so`` reached through
    the C ABI in ``include/oneprot_clip.h`` (tcgen05/TMEM/TMA for the contractions);
  * PyTorch supplies device memory, the current stream, autograd bookkeeping and
    ``torch.distributed`` collectives (all-gather of the second operand, a 3N-float all-reduce of
    the softmax sums, reduce-scatter of the partial gradient) - the reference's own exchange steps
    (loss.py:32-38) without its W-fold redundant N x N compute (loss.py:95).
There is no CPU or eager fallback: without the CUDA library the forward raises.

Sharding: rank r owns rows [r n, (r+1) n) of the logit matrix.  It computes
the row panel Z[r, :] tile by tile, which gives its row sums exactly and partial column sums;
one small all-reduce completes the column sums.  Backward recomputes the panel into a bounded
bf16 dL/dZ workspace (``panel_bytes``), dA_r is complete locally and the partial dB is
reduce-scattered.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn

from . import kernels as _cuda_kernels
from . import sequencer as _seq

try:  # same guard as the reference (loss.py:5-11)
    import torch.distributed as dist
    has_distributed = dist.is_available()
except ImportError:  # pragma: no cover
    dist = None
    has_distributed = False

# Kernel provider.  Tests of the host-side sharding logic (CPU, gloo) swap this for an emulation
# that lives under tests/; the product never does.
_KERNELS = _cuda_kernels

# bound of the bf16 dL/dZ panel workspace: 1.25 GiB lets N = 32768 run as two wave-aligned panels
DEFAULT_PANEL_BYTES = 5 << 28

# bound of the stored-exponentials panel (keep_exp=True): n x N bf16 per rank, 2 GiB at N = 32768 on one GPU
DEFAULT_KEEP_BYTES = 8 << 30

_SCALE_CACHE = {}               # (device, python float) -> 1-element fp32 device tensor


def _float_scale_on(device, value: float) -> torch.Tensor:
    key = (str(device), float(value))
    t = _SCALE_CACHE.get(key)
    if t is None:
        if len(_SCALE_CACHE) > 64:
            _SCALE_CACHE.clear()
        t = torch.full((1,), float(value), dtype=torch.float32, device=device)
        _SCALE_CACHE[key] = t
    return t


# ---------------------------------------------------------------------------------------------
# collectives (plumbing only)
# ---------------------------------------------------------------------------------------------
from .comm import _all_gather_rows, _reduce_scatter_rows, make_comm   # noqa: E402

_COMMS = {}     # (device, group id, world, rank, kernel provider) -> exchange provider (owns the NVLS workspace)


def _get_comm(world, rank, group, device, prefer=None):
    key = (str(device), id(group), world, rank, id(_KERNELS), prefer)
    c = _COMMS.get(key)
    if c is None:
        c = make_comm(_KERNELS, world, rank, group, device, prefer=prefer)
        _COMMS[key] = c
    return c


class _AllGatherWithGrad(torch.autograd.Function):
    """all-gather whose backward is a reduce-scatter SUM (``torch.distributed.nn.all_gather``)."""

    @staticmethod
    def forward(ctx, x, rank, world_size, group):
        ctx.rank, ctx.world_size, ctx.group = rank, world_size, group
        return _all_gather_rows(x, world_size, group)

    @staticmethod
    def backward(ctx, g):
        return _reduce_scatter_rows(g, ctx.rank, ctx.world_size, ctx.group), None, None, None


class _InsertLocalRows(torch.autograd.Function):
    """Gathered tensor (no grad) whose rows of this rank are the grad-carrying local tensor -
    the ``gathered[rank] = features`` of loss.py:39-42 without the list + cat."""

    @staticmethod
    def forward(ctx, gathered, local, rank):
        ctx.rank, ctx.n = rank, local.shape[0]
        return gathered

    @staticmethod
    def backward(ctx, g):
        return None, g[ctx.rank * ctx.n:(ctx.rank + 1) * ctx.n], None


def gather_features(modality_features, sequence_features, local_loss=False, gather_with_grad=False, rank=0,
                    world_size=1, use_horovod=False):
    """Same contract as the reference ``gather_features`` (loss.py:19-46): returns the two feature
    tensors concatenated over ranks.  ``use_horovod`` is accepted and ignored, as in the reference."""
    assert has_distributed, 'torch.distributed did not import correctly, please use a PyTorch version with support.'
    if gather_with_grad:
        all_m = _AllGatherWithGrad.apply(modality_features, rank, world_size, None)
        all_s = _AllGatherWithGrad.apply(sequence_features, rank, world_size, None)
    else:
        with torch.no_grad():
            all_m = _all_gather_rows(modality_features, world_size)
            all_s = _all_gather_rows(sequence_features, world_size)
        if not local_loss:
            all_m = _InsertLocalRows.apply(all_m, modality_features, rank)
            all_s = _InsertLocalRows.apply(all_s, sequence_features, rank)
    return all_m, all_s


# ---------------------------------------------------------------------------------------------
# operand preparation
# ---------------------------------------------------------------------------------------------
def _prep_side(x: torch.Tensor, side: int):
    """bf16 GEMM operand of one feature tensor -> (operand, d_padded, split).

    bf16 inputs are used as they are.  fp32 / fp16 inputs are split into bf16 limbs x = h + m (+ l)
    laid out along K so that one bf16 GEMM reproduces the fp32 dot product:
    left (side 0) operand [h | h | m], right (side 1) operand [h | m | h] => <l, r> = hh + hm + mh
    (error 2^-16 per product instead of 2^-8).  Feature dims that are not a multiple of 8 are
    zero-padded (TMA needs 16-byte row pitches; zero columns change no dot product)."""
    x = x.detach()
    d = x.shape[1]
    pad = (-d) % 8
    if pad:
        x = torch.nn.functional.pad(x, (0, pad))
        d += pad
    if x.dtype == torch.bfloat16:
        return x.contiguous(), d, False
    out = torch.empty(x.shape[0], 3 * d, dtype=torch.bfloat16, device=x.device)
    _KERNELS.split_fp32(x.float().contiguous(), out, side, 3)
    return out, d, True


class _Operands:
    """The prepared (A, B) pair of one ClipLoss call."""

    def __init__(self, A: torch.Tensor, B: torch.Tensor):
        self.in_dtype = A.dtype
        self.n, self.d_in = A.shape
        self.A, self.d, self.split = _prep_side(A, 0)
        self.B, _, _ = _prep_side(B, 1)
        self.pad = self.d - self.d_in
        self.dk = self.A.shape[1]          # contraction length seen by the kernels

    # column blocks whose sum reconstructs the fp32 operand (for the gradient GEMMs)
    def b_pieces(self, B_all):
        d = self.d
        return [B_all[:, 0:d], B_all[:, d:2 * d]] if self.split else [B_all]

    def a_pieces(self, A_rows):
        d = self.d
        return [A_rows[:, d:2 * d], A_rows[:, 2 * d:3 * d]] if self.split else [A_rows]

    def a_head(self, A_rows):
        return A_rows[:, 0:self.d] if self.split else A_rows


class _GemmChain:
    """Sum of several GEMM terms into one output: fp32 accumulator chained through acc_in /
    acc_out, the last term applies the row scale and writes the final dtype."""

    def __init__(self, M, Nc, out: torch.Tensor, row_scale, n_terms: int):
        self.M, self.Nc, self.out, self.row_scale, self.n_terms = M, Nc, out, row_scale, n_terms
        self.done = 0
        self.acc = None
        if n_terms > 1 and out.dtype != torch.float32:
            self.acc = torch.empty(M, out.stride(0), dtype=torch.float32, device=out.device)[:, :Nc]

    def add(self, A, a_mn, B, b_mn, K, dot_mat=None, rowdot_part=None):
        Kn = _KERNELS
        first, last = self.done == 0, self.done == self.n_terms - 1
        f32_out = self.out.dtype == torch.float32
        acc = self.out if f32_out else self.acc
        kw = dict(dot_mat=dot_mat, rowdot_part=rowdot_part)
        if last:
            kw["row_scale"] = self.row_scale
            if f32_out:
                Kn.gemm_bf16(A, a_mn, B, b_mn, self.M, self.Nc, K, acc_in=None if first else acc, acc_out=self.out, **kw)
            else:
                Kn.gemm_bf16(A, a_mn, B, b_mn, self.M, self.Nc, K, acc_in=None if first else acc, out=self.out, **kw)
        else:
            Kn.gemm_bf16(A, a_mn, B, b_mn, self.M, self.Nc, K, acc_in=None if first else acc, acc_out=acc, **kw)
        self.done += 1


def _robust_forward(K, comm, ops, B_all, scale_dev, diag_loc, n, N, off, W, mode):
    """Forward of the two-reference path.  References: rho_i = max_j x_ij (this rank's rows),
    gamma_j = max_i x_ij (all ranks).  Augmenting the operands with the reference as one more K
    column makes the unchanged tensor-core kernels produce x_ij - rho'_i (row direction) resp.
    x_ij - gamma'_j (column direction) with G = 0, so every sum is in [1, N]: no overflow, no
    underflow, for any inputs.  Costs one max pass + two sum passes (3 GEMM units instead of 1)."""
    dev = ops.A.device
    dk = ops.A.shape[1]
    rowmax = torch.empty(n, dtype=torch.float32, device=dev)
    colmax = torch.empty(N, dtype=torch.float32, device=dev)
    K.rowcol_max(ops.A, B_all, scale_dev, rowmax, colmax)
    if W > 1:
        dist.all_reduce(colmax, op=dist.ReduceOp.MAX, group=comm.group)
    A_r = torch.empty(n, dk + 8, dtype=torch.bfloat16, device=dev)
    A_1 = torch.empty(n, dk + 8, dtype=torch.bfloat16, device=dev)
    B_1 = torch.empty(N, dk + 8, dtype=torch.bfloat16, device=dev)
    B_c = torch.empty(N, dk + 8, dtype=torch.bfloat16, device=dev)
    # vec: [colsum | rowsum | diag | row_ref] in global order, summed over ranks
    vec = torch.zeros(4 * N, dtype=torch.float32, device=dev)
    col_ref = torch.empty(N, dtype=torch.float32, device=dev)
    K.augment(ops.A, rowmax, scale_dev, A_r, vec[3 * N + off:3 * N + off + n])
    K.augment(B_all, None, scale_dev, B_1, None)
    K.augment(ops.A, None, scale_dev, A_1, None)
    K.augment(B_all, colmax, scale_dev, B_c, col_ref)
    stats = torch.zeros(4, dtype=torch.float32, device=dev)
    stats[3] = 1.0                                   # exact-reference override with stats[2] = 0  =>  G = 0
    junk_r = torch.empty(n, dtype=torch.float32, device=dev)
    junk_c = torch.empty(N, dtype=torch.float32, device=dev)
    K.fwd_sums(A_r, B_1, scale_dev, stats, vec[N + off:N + off + n], junk_c)      # row sums (complete)
    K.fwd_sums(A_1, B_c, scale_dev, stats, junk_r, vec[0:N])                      # column sums (partial)
    vec[2 * N + off:2 * N + off + n].copy_(diag_loc)
    if W > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=comm.group)
    out = torch.empty(1 + 2 * N, dtype=torch.float32, device=dev)
    loss32, inv_rs, inv_cs = out[0:1], out[1:1 + N], out[1 + N:]
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    K.loss_finalize(vec[N:2 * N], vec[0:N], vec[2 * N:3 * N], n, off, mode, scale_dev, stats, loss32, inv_rs, inv_cs, flag,
                    row_ref=vec[3 * N:4 * N], col_ref=col_ref)
    return dict(loss32=loss32, inv_rs=inv_rs, inv_cs=inv_cs, flag=flag, stats=stats,
                dz_ops={1: (A_r, B_1), 2: (A_1, B_c)})


# ---------------------------------------------------------------------------------------------
# the autograd function
# ---------------------------------------------------------------------------------------------
class _ClipLossFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, A, B, scale_t, cfg):
        with _KERNELS.stream_scope():
            return _ClipLossFunction._forward_impl(ctx, A, B, scale_t, cfg)

    @staticmethod
    def backward(ctx, g_loss, _g32, _gflag):
        with _KERNELS.stream_scope():
            return _ClipLossFunction._backward_impl(ctx, g_loss)

    @staticmethod
    def _forward_impl(ctx, A, B, scale_t, cfg):
        K = _KERNELS
        W, rank, group = cfg["world_size"], cfg["rank"], cfg["group"]
        dev = A.device
        ops = _Operands(A, B)
        n, off = ops.n, rank * ops.n
        N = W * n
        comm = _get_comm(W, rank, group, dev, prefer=cfg.get("comm_prefer"))
        scale_dev = scale_t.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        loss_dtype = cfg["loss_dtype"] or ops.in_dtype

        # Stored-exponentials backward: the forward keeps e_ij as a bf16 n x N panel and the backward
        # rescales it in place instead of recomputing the logits (3 GEMM units per step instead of 4).  Needs the
        # one-pass gradient conventions, bf16 operands, the single-reference path and the panel to fit keep_bytes.
        local_mode = W > 1 and cfg["local_loss"]
        ldw = (N + 63) // 64 * 64
        keep = None
        if (cfg.get("keep_exp") and not ops.split and cfg.get("robust", "off") == "off"
                and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or scale_t.requires_grad)
                and not (local_mode and (not cfg["gather_with_grad"] or scale_t.requires_grad))
                and 2 * ldw * ((n + 127) // 128 * 128) <= cfg.get("keep_bytes", DEFAULT_KEEP_BYTES)):
            keep = torch.empty((n + 127) // 128 * 128, ldw, dtype=torch.bfloat16, device=dev)

        if K is _cuda_kernels and _seq.enabled(cfg) and _seq.eligible(cfg, ops, comm, scale_t.requires_grad):
            # host-side step sequencer: the launches below, issued from one C call per phase
            loss32, flag = _seq.forward(ctx, ops, scale_dev, cfg, comm, K, keep=keep)
            ctx.cfg, ctx.ops, ctx.comm = cfg, ops, comm
            ctx.set_materialize_grads(False)
            loss_out = loss32.reshape(()).to(loss_dtype)
            loss_f32 = loss32.reshape(()).clone()
            ctx.mark_non_differentiable(loss_f32, flag)
            return loss_out, loss_f32, flag

        st = comm.begin_forward(ops, rank, W)          # second operand of all ranks + zeroed [colsum | rowsum | diag]
        B_all, sums = st["B_all"], st["sums"]
        K.rowstats(ops.A, st["stats_rows"], st["stats_off"], sums[2 * N + off:2 * N + off + n], st["stats"])
        stats = comm.global_stats(st)                  # one reference G on every rank
        if st.get("ag") is not None:     # all-gather fused into the forward kernel (NVLS provider)
            K.fwd_sums(ops.A, B_all, scale_dev, stats, sums[N + off:N + off + n], sums[0:N], ag=st["ag"],
                       **({"keep": keep} if keep is not None else {}))
            stats = st["stats_glob"]     # global maxima, written by the kernel
        else:
            K.fwd_sums(ops.A, B_all, scale_dev, stats, sums[N + off:N + off + n], sums[0:N],
                       **({"keep": keep} if keep is not None else {}))
        sums = comm.complete_sums(st)
        colsum, rowsum_all, diag_all = sums[0:N], sums[N:2 * N], sums[2 * N:3 * N]

        out = torch.empty(1 + 2 * N, dtype=torch.float32, device=dev)
        loss32, inv_rs, inv_cs = out[0:1], out[1:1 + N], out[1 + N:]
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        mode = K.MODE_LOCAL if (W > 1 and cfg["local_loss"]) else K.MODE_GLOBAL
        K.loss_finalize(rowsum_all, colsum, diag_all, n, off, mode, scale_dev, stats, loss32, inv_rs, inv_cs, flag)

        ctx.dz_ops = None
        ctx.E = keep                    # consumed (overwritten) by the first backward
        robust = cfg.get("robust", "off")
        if robust == "always" or (robust == "auto" and int(flag.item()) != 0):     # "auto" pays one host sync
            # Two-reference path: per-row and per-column references, one pass per softmax direction
            # on operands augmented with the reference as an extra K column.
            B_all = comm.b_all_for_backward(ops, B_all, st["token"], rank, W)
            if W > 1 and comm.name == "nvls":
                B_all = B_all.clone()            # the symmetric buffer is recycled by later forwards
            comm = _get_comm(W, rank, group, dev, prefer="dist") if W > 1 else comm
            rr = _robust_forward(K, comm, ops, B_all, scale_dev, diag_all[off:off + n], n, N, off, W, mode)
            loss32, inv_rs, inv_cs, flag, stats = rr["loss32"], rr["inv_rs"], rr["inv_cs"], rr["flag"], rr["stats"]
            ctx.dz_ops = rr["dz_ops"]
            st = dict(st, token=None)

        ctx.cfg, ctx.ops, ctx.mode, ctx.comm, ctx.token = cfg, ops, mode, comm, st["token"]
        needs_bwd = ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or scale_t.requires_grad
        ctx.bhold = comm.hold_for_backward(B_all) if (needs_bwd and st["token"] is not None) else None
        ctx.B_all, ctx.scale_dev, ctx.stats, ctx.inv_rs, ctx.inv_cs = B_all, scale_dev, stats, inv_rs, inv_cs
        ctx.scale_needs_grad = scale_t.requires_grad
        ctx.scale_meta = (scale_t.dtype, scale_t.device, scale_t.shape)
        ctx.set_materialize_grads(False)
        loss_out = loss32.reshape(()).to(loss_dtype)
        loss_f32 = loss32.reshape(()).clone()
        ctx.mark_non_differentiable(loss_f32, flag)
        return loss_out, loss_f32, flag

    @staticmethod
    def _backward_impl(ctx, g_loss):
        K = _KERNELS
        if getattr(ctx, "seq", None) is not None:
            dA, dB = _seq.backward(ctx, g_loss, ctx.cfg, ctx.ops, ctx.comm, K)
            return (_finish_grad(dA, ctx.ops, ctx.needs_input_grad[0]), _finish_grad(dB, ctx.ops, ctx.needs_input_grad[1]),
                    None, None)
        cfg, ops, mode = ctx.cfg, ctx.ops, ctx.mode
        W, rank, group = cfg["world_size"], cfg["rank"], cfg["group"]
        n, d, dk = ops.n, ops.d, ops.dk
        N, off = W * n, rank * n
        dev = ops.A.device
        comm = ctx.comm
        B_all = comm.b_all_for_backward(ops, ctx.B_all, ctx.token, rank, W, hold=getattr(ctx, "bhold", None))
        need_a, need_b, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.scale_needs_grad

        g32 = (torch.zeros(1, dtype=torch.float32, device=dev) if g_loss is None
               else g_loss.detach().to(device=dev, dtype=torch.float32).reshape(1))
        local = mode == K.MODE_LOCAL
        gwg = cfg["gather_with_grad"]
        # Overlap (NVLS provider): in the global modes the panel is built for a unit upstream gradient
        # and g only scales the GEMM outputs, so its exchange runs on a side stream under the dL/dZ
        # kernel; the pull-reduce of the partial dB runs on the same side stream under the dA GEMM.
        side = comm.side_stream(dev) if W > 1 else None
        main = torch.cuda.current_stream() if side is not None else None
        ev_g = None
        if side is not None and not local:
            gvec = torch.zeros((W + 3) // 4 * 4, dtype=torch.float32, device=dev)   # placeholder for `what=1`
            g_holder = {}
            ev0 = main.record_event()
            with torch.cuda.stream(side), K.stream_scope():
                side.wait_event(ev0)
                g_holder["gvec"] = comm.gather_grad_outputs(g32, rank, W, ctx.token)
        else:
            gvec = comm.gather_grad_outputs(g32, rank, W, ctx.token)
            g_holder = None
        # cross-rank gradient only flows where the reference's graph has it
        exchange_b = W > 1
        if local and not gwg:
            passes = [(1, need_a or need_s, False), (2, False, need_b or need_s)]
        elif (local and need_s) or ctx.dz_ops is not None:
            passes = [(1, True, True), (2, True, True)]       # one pass per softmax direction
        else:
            passes = [(0, need_a or need_s, need_b or (W > 1))]
        # collectives must be entered by every rank: with W > 1 the partial dB is always exchanged
        grad_dtype = torch.float32 if ops.split else torch.bfloat16
        ldw = (N + 63) // 64 * 64
        rows_cap = max(128, (cfg["panel_bytes"] // (2 * ldw)) // 128 * 128)
        if rows_cap < n:
            # several panels: balance them and make the dA GEMM of every full panel an integer number
            # of waves (its 128 x 256 tiles are dealt to one persistent CTA per SM)
            unit = K.panel_row_unit(d)
            n_panels = -(-n // rows_cap)
            target = -(-n // n_panels)
            if unit <= rows_cap:
                up = -(-target // unit) * unit
                rows_cap = up if up <= rows_cap else rows_cap // unit * unit
            else:
                rows_cap = min(rows_cap, -(-target // 128) * 128)
        panels = [(r0, min(rows_cap, n - r0)) for r0 in range(0, n, rows_cap)]
        E = getattr(ctx, "E", None)
        if E is not None and len(passes) == 1:
            # stored exponentials: the whole n x N panel exists already; one in-place rescale replaces the dL/dZ
            # recompute.  A second backward over the same graph (retain_graph) finds E consumed and recomputes.
            # (Rescaling panel by panel on a second stream under the GEMMs was measured in round 2: 6.20 - 6.68 ms
            # per step against 6.14 ms for the single pass at N = 32768 - removed.)
            ctx.E = None
            Wz = E
            panels = [(0, n)]
        else:
            E = None
            Wz = torch.empty(min(rows_cap, (n + 127) // 128 * 128), ldw, dtype=torch.bfloat16, device=dev)

        vec = torch.empty(3 * n + 2 * N, dtype=torch.float32, device=dev)
        wr, dg, sA = vec[0:n], vec[n:2 * n], vec[2 * n:3 * n]
        wc, sB = vec[3 * n:3 * n + N], vec[3 * n + N:]

        dA_total = dB_total = None
        ds_terms = []      # 1-element tensors whose sum is scale * d(value)/d(scale) * g  (this rank's part)
        b_pieces, n_bp = ops.b_pieces(B_all), (2 if ops.split else 1)
        for pi, (part, want_a, want_b) in enumerate(passes):
            if g_holder is not None:
                K.bwd_weights(ctx.inv_rs, ctx.inv_cs, n, off, mode, gwg, part, W, rank, gvec, ctx.scale_dev,
                              wr, wc, dg, sA, sB, 1)                     # panel weights: no g needed
                with torch.cuda.stream(side), K.stream_scope():
                    K.bwd_weights(ctx.inv_rs, ctx.inv_cs, n, off, mode, gwg, part, W, rank, g_holder["gvec"],
                                  ctx.scale_dev, wr, wc, dg, sA, sB, 2)  # output scales from the gathered g
                    ev_g = side.record_event()
            else:
                K.bwd_weights(ctx.inv_rs, ctx.inv_cs, n, off, mode, gwg, part, W, rank, gvec, ctx.scale_dev,
                              wr, wc, dg, sA, sB)
            dA = torch.empty(n, d, dtype=grad_dtype, device=dev) if want_a else None
            dBp = comm.db_buffer(N, d, grad_dtype, dev) if want_b else None
            chain_b = _GemmChain(N, d, dBp, sB, len(panels) * n_bp) if want_b else None
            rd_global = need_s and not local         # rowdot of the unscaled dA rows inside the GEMM epilogue
            last_pass = pi == len(passes) - 1
            ev_rs, dB_async = None, None

            def run_dA(r0, rows, Wp, A_rows):
                chain_a = _GemmChain(rows, d, dA[r0:r0 + rows], sA[r0:r0 + rows], n_bp)
                for bi, Bp in enumerate(b_pieces):
                    if rd_global and bi == n_bp - 1:   # last piece: the accumulated (unscaled) value
                        part_buf = torch.empty(K.gemm_rowdot_scratch_floats(rows, d), dtype=torch.float32, device=dev)
                        chain_a.add(Wp, False, Bp, True, N, dot_mat=ops.a_head(A_rows), rowdot_part=part_buf)
                        t = torch.empty(1, dtype=torch.float32, device=dev)
                        # rows >= `rows` of a slab are never written: sum only the valid part
                        K.sum_f32(_valid_rowdot(part_buf, rows, d), t)
                        ds_terms.append(("unit", t))
                    else:
                        chain_a.add(Wp, False, Bp, True, N)

            for qi, (r0, rows) in enumerate(panels):
                A_rows = ops.A[r0:r0 + rows]
                if E is not None:
                    K.dz_from_exp(E, n, N, off, wr, wc, dg)
                elif ctx.dz_ops is not None:     # two-reference path: augmented operands of this direction
                    A_aug, B_aug = ctx.dz_ops[part]
                    K.dz_panel(A_aug[r0:r0 + rows], B_aug, off + r0, ctx.scale_dev, ctx.stats, wr[r0:r0 + rows], wc,
                               dg[r0:r0 + rows], Wz)
                else:
                    K.dz_panel(A_rows, B_all, off + r0, ctx.scale_dev, ctx.stats, wr[r0:r0 + rows], wc, dg[r0:r0 + rows], Wz)
                Wp = Wz[r0:r0 + rows] if E is not None else Wz[:rows]
                if ev_g is not None:
                    main.wait_event(ev_g)        # GEMM epilogues read the output scales
                    ev_g = None
                if want_b:                        # dB first: its exchange then hides under the dA GEMM
                    for Ap in ops.a_pieces(A_rows):
                        chain_b.add(Wp, True, Ap, True, rows)
                    if exchange_b and side is not None and qi == len(panels) - 1:
                        evb = main.record_event()
                        with torch.cuda.stream(side), K.stream_scope():
                            side.wait_event(evb)
                            dB_async = comm.reduce_scatter_db(dBp, rank, W, last_pass=last_pass)
                            ev_rs = side.record_event()
                if want_a:
                    run_dA(r0, rows, Wp, A_rows)
            if want_b and exchange_b:
                if dB_async is not None:
                    main.wait_event(ev_rs)
                    dB_async.record_stream(main)
                    dBp = dB_async
                else:
                    dBp = comm.reduce_scatter_db(dBp, rank, W, last_pass=last_pass)
            if local and need_s:
                # local loss: d value_r / d scale = (sum_i <a_i, dA^P_i> + sum_j <b_j, dB^Q_j>) / scale
                if part == 1 and dA is not None:
                    ds_terms.append(("scaled", _rowdot_sum(ops.a_head(ops.A), dA)))
                if part == 2 and dBp is not None:
                    ds_terms.append(("scaled", _rowdot_sum(ops.B[:, 0:d], dBp)))
            if local and not gwg:
                dA_total = dA if part == 1 else dA_total
                dB_total = dBp if part == 2 else dB_total
            else:
                dA_total = dA if dA_total is None else (dA_total + dA if dA is not None else dA_total)
                dB_total = dBp if dB_total is None else (dB_total + dBp if dBp is not None else dB_total)

        grad_a = _finish_grad(dA_total, ops, need_a)
        grad_b = _finish_grad(dB_total, ops, need_b)
        grad_s = None
        if need_s:
            tot = torch.zeros(1, dtype=torch.float32, device=dev)
            for kind, t in ds_terms:
                tot = tot + t
            if not local:
                # unit-gradient partial of this rank's rows -> all ranks' rows, times this rank's upstream g
                tot = comm.sum_scalar(tot)
                tot = tot * g32
            grad_s = tot / ctx.scale_dev
            sdt, sdev, sshape = ctx.scale_meta
            grad_s = grad_s.reshape(sshape).to(device=sdev, dtype=sdt)
        return grad_a, grad_b, grad_s, None


def _valid_rowdot(part_buf, rows, d):
    """The written entries of the rowdot partial buffer: [slabs, ldd] -> [:, :rows]."""
    ldd = (rows + 127) // 128 * 128
    slabs = part_buf.numel() // ldd
    return part_buf.view(slabs, ldd)[:, :rows].contiguous()


def _rowdot_sum(X, G):
    """sum_i <x_i, g_i> with the rowdot + sum kernels (G in bf16 or fp32 -> bf16 view for the dot)."""
    K = _KERNELS
    Gb = G if G.dtype == torch.bfloat16 else G.to(torch.bfloat16)
    rows, d = Gb.shape
    tmp = torch.empty(rows, dtype=torch.float32, device=G.device)
    K.rowdot_bf16(X[:, :d] if X.shape[1] != d else X, Gb, tmp)
    t = torch.empty(1, dtype=torch.float32, device=G.device)
    K.sum_f32(tmp, t)
    return t


def _finish_grad(g, ops, needed):
    if not needed or g is None:
        return None
    if ops.pad:
        g = g[:, :ops.d_in]
    return g.to(ops.in_dtype) if g.dtype != ops.in_dtype else g


# ---------------------------------------------------------------------------------------------
# the module
# ---------------------------------------------------------------------------------------------
class ClipLoss(nn.Module):
    """B200-native drop-in for the reference ``ClipLoss`` (loss.py:49-114).

    Extra keyword-only arguments (all optional, defaults keep reference behaviour):
      loss_dtype   dtype of the returned scalar; default = input dtype like the reference
                   (bf16 in -> bf16 out).  The fp32 value is always kept in ``last_loss_fp32``.
      panel_bytes  bound of the bf16 dL/dZ panel workspace used by the backward.
      group        process group for the collectives (default: the world group).
      host_sequencer  (default True) enqueue each phase of the step from one C call (sequencer.py, csrc/clip_sequence.cu)
                   instead of kernel by kernel from Python wherever that path covers the call (bf16 features, logit_scale
                   without gradient, one-pass conventions, NVLS exchange); the results are bit-identical, the host cost
                   of a small step drops by a third (B200, 5 pairs x 1024 rows: 1.10 vs 1.78 ms on one GPU, 1.81 vs
                   3.16 ms on two).  False forces the Python host.
      graph        replay the forward / backward launch sequences as CUDA graphs; removes most of the host enqueue
                   that dominates at OneProt's batch sizes (5 pairs x 1024 rows on one B200: 0.77 ms against 1.10 ms
                   with the C sequencer).  With world_size > 1 the step is captured over a STATIC variant of the NVLS
                   exchange provider (single buffer set, separate gather kernel, 5 symmetric-memory barriers per step;
                   needs the NVLS provider).  A backward must follow its own forward (a stale one raises).
      keep_exp     stored-exponentials backward (default True): the forward keeps the n x N exponentials as a
                   bf16 panel (<= keep_bytes, default 8 GiB; 2 GiB at N = 32768 on one GPU) and the backward rescales
                   it in place instead of recomputing the logits - 3 GEMM units per step instead of 4 (measured on
                   B200, N = 32768: 6.14 vs 7.07 ms per step).  False, a panel above keep_bytes, fp32 features, the
                   two-reference path and the two-pass gradient conventions use the recompute backward.
      check_rows   "first" (default): the first call with a new (n, d, dtype) gathers n, d of every rank (one host
                   sync, cached per shape) and raises ValueError when the ranks disagree - the reference leaves that
                   to a collective error or hang (oneprot_datamodule.py:72 drop_last=False makes it possible);
                   "always": every call; "never": skip.
      robust       "off": one common reference, validated window, device flag;
                   "always": per-row / per-column references for arbitrary inputs (2.25x the work);
                   "auto": run the normal path, read the device flag (one host sync per forward)
                   and fall back to the two-reference path only when it is raised.
                   Default (None): "off" while gradients are enabled (training: no host sync), "auto"
                   under torch.no_grad() - the reference's test_step scales the already scaled modality
                   features a second time (oneprot_module.py:142), which spreads the logits
                   over +-200 nats, beyond what one common reference can hold in fp32.
    """

    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1,
                 use_horovod=False, *, loss_dtype: Optional[torch.dtype] = None,
                 panel_bytes: int = DEFAULT_PANEL_BYTES, group=None, host_sequencer: bool = True,
                 robust: Optional[str] = None, graph: bool = False, keep_exp: bool = True,
                 keep_bytes: int = DEFAULT_KEEP_BYTES, check_rows: str = "first"):
        super().__init__()
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        self.use_horovod = use_horovod
        self.loss_dtype = loss_dtype
        self.panel_bytes = int(panel_bytes)
        self.group = group
        self.host_sequencer = bool(host_sequencer)
        if robust not in (None, "off", "auto", "always"):
            raise ValueError("robust must be None, 'off', 'auto' or 'always'")
        self.robust = robust
        if graph and robust == "auto":
            raise ValueError("graph=True needs a robust mode that is decided on the host side up front")
        self.graph = bool(graph)
        self.keep_exp = bool(keep_exp)
        self.keep_bytes = int(keep_bytes)
        if check_rows not in ("first", "always", "never"):
            raise ValueError("check_rows must be 'first', 'always' or 'never'")
        self.check_rows = check_rows
        self._rows_checked = set()   # (n, d, dtype) already compared across ranks
        self._graphs = {}            # (shape, dtype, gradient pattern, scale kind) -> GraphedStep
        # cache state (same attributes as the reference, loss.py:68-70)
        self.prev_num_logits = 0
        self.labels = {}
        # diagnostics of the last call (device tensors, no host sync)
        self.last_loss_fp32 = None
        self.last_hazard_flag = None

    def get_ground_truth(self, device, num_logits) -> torch.Tensor:
        """Labels of the reference (loss.py:72-83).  The fused kernels use the diagonal implicitly;
        this stays for API parity."""
        if self.prev_num_logits != num_logits or device not in self.labels:
            labels = torch.arange(num_logits, device=device, dtype=torch.long)
            if self.world_size > 1 and self.local_loss:
                labels = labels + num_logits * self.rank
            if self.cache_labels:
                self.labels[device] = labels
                self.prev_num_logits = num_logits
        else:
            labels = self.labels[device]
        return labels

    def get_logits(self, modality_features, sequence_features, logit_scale):
        """Materialised logits as in loss.py:85-101 - DEBUG / API parity only (the loss path never forms them).
        Like the reference, the scale is multiplied into the LEFT operand and rounded to the feature dtype before the
        contraction (loss.py:92-99); the contraction runs on the tcgen05 GEMM kernel with fp32
        accumulators and the result is rounded to the feature dtype (bf16 in -> bf16 out, as torch's matmul).
        The three cases of the reference: world 1 -> (s A) B^T and (s B) A^T; sharded global -> Z and Z.T on the
        gathered operands; sharded local -> this rank's rows against the gathered other side."""
        K = _KERNELS
        A, B = modality_features, sequence_features
        if self.world_size > 1:
            A_all, B_all = gather_features(A, B, self.local_loss, self.gather_with_grad, self.rank, self.world_size,
                                           self.use_horovod)

        def mm(x, y):
            xs = (logit_scale * x).to(x.dtype)           # scale-then-round, the reference's operator precedence
            xo, _, _ = _prep_side(xs, 0)
            yo, _, _ = _prep_side(y, 1)
            M, Nc, Kd = xo.shape[0], yo.shape[0], xo.shape[1]
            ld = (Nc + 7) // 8 * 8                   # the GEMM writes whole 8-column groups: zero rows of y pad the tail
            if ld != Nc:
                yo = torch.nn.functional.pad(yo, (0, 0, 0, ld - Nc))
            out = torch.empty(M, ld, dtype=torch.float32, device=x.device)
            K.gemm_bf16(xo, False, yo, False, M, ld, Kd, acc_out=out)
            return out[:, :Nc].to(x.dtype)

        with torch.no_grad():
            if self.world_size > 1 and self.local_loss:
                return mm(A, B_all), mm(B, A_all)
            if self.world_size > 1:
                z = mm(A_all, B_all)
                return z, z.T
            return mm(A, B), mm(B, A)

    def forward(self, modality_features, sequence_features, logit_scale=1.0, output_dict=False):
        A, B = modality_features, sequence_features
        if A.dim() != 2 or B.dim() != 2 or A.shape != B.shape:
            raise ValueError(f"ClipLoss expects two (n, d) tensors of equal shape, got {tuple(A.shape)} and {tuple(B.shape)}")
        if A.dtype != B.dtype or A.device != B.device:
            raise ValueError("ClipLoss expects both feature tensors on one device with one dtype")
        if A.dtype not in (torch.bfloat16, torch.float32, torch.float16):
            raise ValueError(f"unsupported feature dtype {A.dtype}")
        if self.world_size > 1:
            assert has_distributed, 'torch.distributed did not import correctly, please use a PyTorch version with support.'
            if not dist.is_initialized():
                raise RuntimeError("ClipLoss(world_size > 1) needs an initialised torch.distributed process group")
            if dist.get_world_size(self.group) != self.world_size:
                raise RuntimeError("ClipLoss world_size does not match the process group")
            self._check_equal_rows(A)
        if torch.is_tensor(logit_scale):
            if logit_scale.numel() != 1:
                raise ValueError("logit_scale must be a scalar")
            scale_t = logit_scale
        else:
            scale_t = _float_scale_on(A.device, logit_scale)   # cached: no host-to-device copy per call
        cfg = dict(world_size=self.world_size, rank=self.rank, group=self.group, local_loss=bool(self.local_loss),
                   gather_with_grad=bool(self.gather_with_grad), loss_dtype=self.loss_dtype,
                   panel_bytes=self.panel_bytes, host_sequencer=self.host_sequencer, keep_exp=self.keep_exp,
                   keep_bytes=self.keep_bytes)
        if self.robust is None:     # training: never sync; evaluation: fall back to the two-reference path when flagged
            cfg["robust"] = "off" if (torch.is_grad_enabled() or self.graph) else "auto"
        else:
            cfg["robust"] = self.robust
        if self.graph and A.is_cuda:
            from .graphed import GraphedClipFunction, GraphedStep
            cfg["host_sequencer"] = False      # launch cost is paid once, at capture: the plain host order is captured
            if self.world_size > 1:      # its own single-buffer workspace per captured shape (addresses are baked into the graphs)
                cfg["comm_prefer"] = f"nvls-static/{A.shape[0]}x{A.shape[1]}/{A.dtype}"
            needs = (bool(A.requires_grad and torch.is_grad_enabled()), bool(B.requires_grad and torch.is_grad_enabled()))
            key = (tuple(A.shape), A.dtype, needs, bool(scale_t.requires_grad), A.device.index)
            step = self._graphs.get(key)
            if step is None:
                # gradients are always captured for tensors that may need them; the eager bodies decide by `needs`
                step = self._graphs[key] = GraphedStep(_ClipLossFunction, A, B, scale_t, cfg, (True, True))
            total_loss, loss32, flag = GraphedClipFunction.apply(A, B, scale_t, step)
        else:
            total_loss, loss32, flag = _ClipLossFunction.apply(A, B, scale_t, cfg)
        self.last_loss_fp32, self.last_hazard_flag = loss32, flag
        return {"contrastive_loss": total_loss} if output_dict else total_loss

    def _check_equal_rows(self, A):
        """Every rank must hold the same (n, d): the row sharding, the symmetric-memory layout and the collectives
        all assume it.  Compared once per new shape (all ranks meet a new shape in the same call when the sampler
        pads ranks to equal length, as DistributedSampler does), on every rank, so all of them raise together."""
        key = (int(A.shape[0]), int(A.shape[1]), A.dtype)
        if self.check_rows == "never" or (self.check_rows == "first" and key in self._rows_checked):
            return
        mine = torch.tensor([key[0], key[1]], dtype=torch.int64, device=A.device)
        every = torch.empty(self.world_size * 2, dtype=torch.int64, device=A.device)
        dist.all_gather_into_tensor(every, mine, group=self.group)
        every = every.view(self.world_size, 2).cpu()
        if not bool((every == every[0]).all()):
            raise ValueError("ClipLoss: every rank must pass the same (n, d); the ranks hold "
                             f"{[tuple(int(v) for v in r) for r in every]} (drop the ragged last batch or pad it)")
        self._rows_checked.add(key)

    def check_last_call(self):
        """Host-side check (synchronises) that the last forward stayed inside the validated fp32
        window of the single-reference exp-sum ('numerical window')."""
        if self.last_hazard_flag is not None and int(self.last_hazard_flag.item()) != 0:
            raise FloatingPointError(
                "ClipLoss: logits left the validated exp2 window (|logit_scale| * max|a| * max|b| is huge and the "
                "row/column maxima are more than ~100 log2 units apart); normalise the features or lower logit_scale")
