"""Drop-in ``SigLipLoss`` for OneProt on B200 - the other objective behind the reference's ``loss_fn``
switch (``src/models/oneprot_module.py:57-62``; reference class ``src/models/components/loss.py:204-311``).

Same constructor ``SigLipLoss(cache_labels, rank, world_size, bidir, use_horovod)`` and
``forward(modality_features, sequence_features, logit_scale=1.0, logit_bias=None, output_dict=False)``.

    value_r = -(1/n) sum_{i in rank r} sum_{j} logsigmoid(y_ij z_ij),   z = logit_scale <a_i, b_j> + logit_bias,
    y = +1 on the global diagonal, -1 elsewhere.

The reference visits the blocks of other ranks by passing the second operand round a ring of
``batch_isend_irecv`` neighbour exchanges (loss.py:258-309), W - 1 hops of one n x n block each.  Here
rank r computes its whole n x N row panel in one pass of the tensor-core mainloop of the ClipLoss
path (``clip_s_kernel<SFWD>``: softplus row sums, nothing stored per logit) on the all-gathered second
operand; the backward recomputes the panel into the bounded bf16 dL/dZ workspace
(``clip_s_kernel<SDZ>``: (g/n)(sigma(z) - [i == j])), dA = Wz B_all is local and the partial
dB = Wz^T A is reduce-scattered - which is what the autograd of the ring (loss.py:169-201) adds up.
``bidir`` only changes the reference's hop schedule, not the result; it is accepted and ignored.

``logit_scale`` / ``logit_bias`` may be Python floats or 0-dim tensors; tensors that require grad get their
gradient like in the reference (d scale = (1/scale) sum_i <a_i, dA_i>, the row-dot of the gradient that is
computed anyway; d bias = (g/n) (sum_ij sigma(z_ij) - n) from row sums the panel kernel adds up on the
side).  As in the reference each rank's scalar gradients are those of ITS loss (DDP all-reduces
parameters later).  Exchanges go through torch.distributed (all-gather / reduce-scatter).  No eager fallback.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import clip_loss as _cl
from .comm import _all_gather_rows, _reduce_scatter_rows

try:
    import torch.distributed as dist
except ImportError:  # pragma: no cover
    dist = None


def _K():
    return _cl._KERNELS          # the kernel provider of the ClipLoss path (tests swap it there)


def _scalar_in(v, device) -> Optional[torch.Tensor]:
    """float -> cached 1-element device tensor (no grad); tensor -> itself (autograd input)."""
    if v is None:
        return None
    if torch.is_tensor(v):
        if v.numel() != 1:
            raise ValueError("logit_scale / logit_bias must be scalars")
        return v
    return _cl._float_scale_on(device, float(v))


def _dev1(t, device):
    return None if t is None else t.detach().to(device=device, dtype=torch.float32).reshape(1).contiguous()


class _SigLipFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, A, B, scale_t, bias_t, cfg):
        with _K().stream_scope():
            K = _K()
            scale_dev, bias_dev = _dev1(scale_t, A.device), _dev1(bias_t, A.device)
            ctx.scalar_meta = [(t.dtype, t.device, t.shape) if t is not None else None for t in (scale_t, bias_t)]
            W, rank, group = cfg["world_size"], cfg["rank"], cfg["group"]
            ops = _cl._Operands(A, B)
            n, off = ops.n, rank * ops.n
            dev = A.device
            # exchange provider of the ClipLoss path (comm.py): on one NVSwitch node the second operand is gathered by one
            # multimem.st pass and the partial dB reduced by multimem.ld_reduce (the reference's ring of W - 1 neighbour
            # exchanges, loss.py:116-201, carries the same rows); torch.distributed collectives otherwise
            ctx.comm = ctx.token = ctx.bhold = None
            if W == 1:
                B_all = ops.B
            else:
                comm = _cl._get_comm(W, rank, group, dev)
                if comm.name == "nvls" and not ops.split:
                    B_all, ctx.token = comm.gather_rows(ops, rank, W)
                    ctx.comm = comm
                    if any(ctx.needs_input_grad[:4]):
                        ctx.bhold = comm.hold_for_backward(B_all)
                else:
                    B_all = _all_gather_rows(ops.B, W, group)
            diag = torch.empty(n, dtype=torch.float32, device=dev)
            stats = torch.zeros(4, dtype=torch.float32, device=dev)
            K.rowstats(ops.A, B_all, off, diag, stats)                  # diag[i] = <a_i, b_{off+i}>: the label logits
            rowsum = torch.empty(n, dtype=torch.float32, device=dev)
            # Kept-panel backward (opt-in): dL/dz_ij = (g / n) (sigma(z_ij) - [i == j]) has a constant weight, so a forward
            # that keeps S = sigma - [i == j] as a bf16 panel makes the backward the two GEMMs on S as it is - no
            # recompute, no rescale pass (3 GEMM units per step instead of 4).
            N = W * n
            ldw = (N + 63) // 64 * 64
            ctx.S = ctx.sig = None
            if (cfg.get("keep_exp") and not ops.split and any(ctx.needs_input_grad[:4])
                    and 2 * ldw * ((n + 127) // 128 * 128) <= cfg.get("keep_bytes", _cl.DEFAULT_KEEP_BYTES)):
                ctx.S = torch.empty((n + 127) // 128 * 128, ldw, dtype=torch.bfloat16, device=dev)
                if bias_dev is not None and ctx.needs_input_grad[3]:
                    ctx.sig = torch.empty(n, dtype=torch.float32, device=dev)
                K.siglip_fwd_keep(ops.A, B_all, off, scale_dev, bias_dev, rowsum, ctx.S, sig_rowsum=ctx.sig)
            else:
                K.siglip_fwd(ops.A, B_all, scale_dev, bias_dev, rowsum)
            loss32 = torch.empty(1, dtype=torch.float32, device=dev)
            K.siglip_finalize(rowsum, diag, scale_dev, bias_dev, loss32)
            ctx.cfg, ctx.ops, ctx.B_all, ctx.scale_dev, ctx.bias_dev = cfg, ops, B_all, scale_dev, bias_dev
            ctx.set_materialize_grads(False)
            loss_f32 = loss32.reshape(()).clone()
            ctx.mark_non_differentiable(loss_f32)
            return loss32.reshape(()).to(cfg["loss_dtype"] or ops.in_dtype), loss_f32

    @staticmethod
    def backward(ctx, g_loss, _g32):
        with _K().stream_scope():
            K = _K()
            cfg, ops, B_all = ctx.cfg, ctx.ops, ctx.B_all
            W, rank, group = cfg["world_size"], cfg["rank"], cfg["group"]
            n, d = ops.n, ops.d
            N, off = W * n, rank * n
            dev = ops.A.device
            if ctx.comm is not None:
                B_all = ctx.comm.b_all_for_backward(ops, B_all, ctx.token, rank, W, hold=ctx.bhold)
            need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
            need_s = ctx.needs_input_grad[2]
            need_bias = ctx.bias_dev is not None and ctx.needs_input_grad[3]
            g32 = (torch.zeros(1, dtype=torch.float32, device=dev) if g_loss is None
                   else g_loss.detach().to(device=dev, dtype=torch.float32).reshape(1))
            # dL/dz_ij = (g / n) (sigma(z_ij) - [i == j]); the panel carries the logit_scale of d z / d <a, b> as well
            coef = (ctx.scale_dev * g32 / n).expand(n).contiguous()
            want_a, want_b = bool(need_a or need_s), bool(need_b or W > 1)   # with W > 1 every rank enters the reduce-scatter
            grad_dtype = torch.float32 if ops.split else torch.bfloat16
            ldw = (N + 63) // 64 * 64
            if ctx.S is not None:
                # kept panel: S already is dL/dz / (logit_scale g / n); the constant goes into the GEMM epilogues' row scale
                S = ctx.S[:n]
                dA = torch.empty(n, d, dtype=grad_dtype, device=dev) if want_a else None
                dBp = _SigLipFunction._db_buffer(ctx, N, d, grad_dtype, dev) if want_b else None
                if want_b:
                    coef_N = (ctx.scale_dev * g32 / n).expand(N).contiguous()
                    _cl._GemmChain(N, d, dBp, coef_N, 1).add(S, True, ops.A, True, n)
                if want_a:
                    _cl._GemmChain(n, d, dA, coef, 1).add(S, False, B_all, True, N)
                return _SigLipFunction._finish(ctx, K, dA, dBp, ctx.sig, g32, need_a, need_b, need_s, need_bias)
            rows_cap = max(128, (cfg["panel_bytes"] // (2 * ldw)) // 128 * 128)
            if rows_cap < n:                                           # balanced, wave-aligned panels (as clip_loss.py)
                unit = K.panel_row_unit(d)
                n_panels = -(-n // rows_cap)
                target = -(-n // n_panels)
                if unit <= rows_cap:
                    up = -(-target // unit) * unit
                    rows_cap = up if up <= rows_cap else rows_cap // unit * unit
                else:
                    rows_cap = min(rows_cap, -(-target // 128) * 128)
            panels = [(r0, min(rows_cap, n - r0)) for r0 in range(0, n, rows_cap)]
            Wz = torch.empty(min(rows_cap, (n + 127) // 128 * 128), ldw, dtype=torch.bfloat16, device=dev)
            n_bp = 2 if ops.split else 1
            dA = torch.empty(n, d, dtype=grad_dtype, device=dev) if want_a else None
            dBp = _SigLipFunction._db_buffer(ctx, N, d, grad_dtype, dev) if want_b else None
            chain_b = _cl._GemmChain(N, d, dBp, None, len(panels) * n_bp) if want_b else None
            b_pieces = ops.b_pieces(B_all)
            sig = torch.empty(n, dtype=torch.float32, device=dev) if need_bias else None     # row sums of sigma(z)
            for r0, rows in panels:
                A_rows = ops.A[r0:r0 + rows]
                K.siglip_dz_panel(A_rows, B_all, off + r0, ctx.scale_dev, ctx.bias_dev, coef[r0:r0 + rows], coef[r0:r0 + rows], Wz,
                                  sig_rowsum=None if sig is None else sig[r0:r0 + rows])
                Wp = Wz[:rows]
                if want_b:
                    for Ap in ops.a_pieces(A_rows):
                        chain_b.add(Wp, True, Ap, True, rows)
                if want_a:
                    chain_a = _cl._GemmChain(rows, d, dA[r0:r0 + rows], None, n_bp)
                    for Bp in b_pieces:
                        chain_a.add(Wp, False, Bp, True, N)
            return _SigLipFunction._finish(ctx, K, dA, dBp, sig, g32, need_a, need_b, need_s, need_bias)

    @staticmethod
    def _db_buffer(ctx, N, d, dtype, dev):
        """Where the partial dB is written: the symmetric workspace under the NVLS provider (its owner pulls the sum)."""
        if ctx.comm is not None and dtype == torch.bfloat16:
            return ctx.comm.db_buffer(N, d, dtype, dev)
        return torch.empty(N, d, dtype=dtype, device=dev)

    @staticmethod
    def _finish(ctx, K, dA, dBp, sig, g32, need_a, need_b, need_s, need_bias):
        """Exchange of the partial dB, scalar gradients, final dtypes (shared by the panel and the kept-panel backward)."""
        cfg, ops = ctx.cfg, ctx.ops
        W, rank, group = cfg["world_size"], cfg["rank"], cfg["group"]
        n, dev = ops.n, ops.A.device
        if dBp is not None and W > 1:
            if ctx.comm is not None and dBp.dtype == torch.bfloat16:
                dBp = ctx.comm.reduce_scatter_db(dBp, rank, W)        # barrier + multimem.ld_reduce pull of this rank's rows
            else:
                dBp = _reduce_scatter_rows(dBp, rank, W, group)
        grad_s = grad_bias = None
        if need_s:      # d loss / d scale = sum_ij dL/dz_ij <a_i, b_j> = (1 / scale) sum_i <a_i, dA_i>
            sdt, sdev, sshape = ctx.scalar_meta[0]
            grad_s = (_cl._rowdot_sum(ops.a_head(ops.A), dA) / ctx.scale_dev).reshape(sshape).to(device=sdev, dtype=sdt)
        if need_bias:   # d loss / d bias = sum_ij dL/dz_ij = (g / n) (sum_ij sigma(z_ij) - n)
            bdt, bdev, bshape = ctx.scalar_meta[1]
            tot = torch.empty(1, dtype=torch.float32, device=dev)
            K.sum_f32(sig, tot)
            grad_bias = ((tot - n) * g32 / n).reshape(bshape).to(device=bdev, dtype=bdt)
        return _cl._finish_grad(dA, ops, need_a), _cl._finish_grad(dBp, ops, need_b), grad_s, grad_bias, None


class SigLipLoss(nn.Module):
    """B200-native drop-in for the reference ``SigLipLoss`` (loss.py:204-311).

    Extra keyword-only arguments: ``loss_dtype`` (dtype of the returned scalar, default = input dtype
    like the reference), ``panel_bytes`` (bound of the bf16 dL/dZ panel), ``group`` (process group),
    ``keep_exp`` / ``keep_bytes`` (default True / 8 GiB: the forward keeps sigma(z) - [i == j] as a bf16 n x N panel and
    the backward is the two GEMMs on it - no recompute, no panel kernel; False or a larger panel: recompute backward)."""

    def __init__(self, cache_labels=False, rank=0, world_size=1, bidir=True, use_horovod=False, *,
                 loss_dtype: Optional[torch.dtype] = None, panel_bytes: int = _cl.DEFAULT_PANEL_BYTES, group=None,
                 keep_exp: bool = True, keep_bytes: int = _cl.DEFAULT_KEEP_BYTES):
        super().__init__()
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        assert not use_horovod  # as the reference (loss.py:229)
        self.use_horovod = use_horovod
        self.bidir = bidir
        self.loss_dtype = loss_dtype
        self.panel_bytes = int(panel_bytes)
        self.group = group
        self.keep_exp = bool(keep_exp)
        self.keep_bytes = int(keep_bytes)
        self.prev_num_logits = 0
        self.labels = {}
        self.last_loss_fp32 = None

    def get_ground_truth(self, device, dtype, num_logits, negative_only=False) -> torch.Tensor:
        """loss.py:237-241 (API parity; the fused kernels use the diagonal implicitly)."""
        labels = -torch.ones((num_logits, num_logits), device=device, dtype=dtype)
        if not negative_only:
            labels = 2 * torch.eye(num_logits, device=device, dtype=dtype) + labels
        return labels

    def get_logits(self, modality_features, sequence_features, logit_scale, logit_bias=None):
        """Materialised logits as in loss.py:243-247 - DEBUG ONLY (tcgen05 GEMM, fp32 accumulators)."""
        z = _cl.ClipLoss(world_size=1).get_logits(modality_features, sequence_features, logit_scale)[0]
        if logit_bias is not None:
            z = z + (logit_bias.to(z.dtype) if torch.is_tensor(logit_bias) else logit_bias)
        return z

    def forward(self, modality_features, sequence_features, logit_scale=1.0, logit_bias=None, output_dict=False):
        A, B = modality_features, sequence_features
        if A.dim() != 2 or B.dim() != 2 or A.shape != B.shape:
            raise ValueError(f"SigLipLoss expects two (n, d) tensors of equal shape, got {tuple(A.shape)} and {tuple(B.shape)}")
        if A.dtype != B.dtype or A.device != B.device:
            raise ValueError("SigLipLoss expects both feature tensors on one device with one dtype")
        if A.dtype not in (torch.bfloat16, torch.float32, torch.float16):
            raise ValueError(f"unsupported feature dtype {A.dtype}")
        if self.world_size > 1:
            if dist is None or not dist.is_initialized():
                raise RuntimeError("SigLipLoss(world_size > 1) needs an initialised torch.distributed process group")
            if dist.get_world_size(self.group) != self.world_size:
                raise RuntimeError("SigLipLoss world_size does not match the process group")
        cfg = dict(world_size=self.world_size, rank=self.rank, group=self.group, loss_dtype=self.loss_dtype,
                   panel_bytes=self.panel_bytes, keep_exp=self.keep_exp, keep_bytes=self.keep_bytes)
        loss, loss32 = _SigLipFunction.apply(A, B, _scalar_in(logit_scale, A.device), _scalar_in(logit_bias, A.device), cfg)
        self.last_loss_fp32 = loss32
        return {"contrastive_loss": loss} if output_dict else loss
