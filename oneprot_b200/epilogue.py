"""L2-normalise / logit-scale epilogue of the OneProt encoders on B200.

Drop-ins for ``Normalize`` and ``LearnableLogitScaling`` of the reference
(``src/models/components/base_encoder.py:6-33``) plus the fused ``NormalizeAndScale`` that the
reference runs as two modules in sequence (``base_encoder.py:171-178``).  Forward and backward are
HBM-bound row kernels of liboneprot_clip.so (one warp per row, 16-byte vector loads, warp-shuffle
reductions); there is no eager fallback.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import kernels as _cuda_kernels

_KERNELS = _cuda_kernels


def _as_rows(x: torch.Tensor):
    """(rows, d) contiguous view/copy in a dtype the kernels take, with d padded to a multiple of 8."""
    d = x.shape[-1]
    x2 = x.reshape(-1, d)
    if x2.dtype not in (torch.bfloat16, torch.float32):
        x2 = x2.float()
    pad = (-d) % 8
    if pad:
        x2 = torch.nn.functional.pad(x2, (0, pad))
    return x2.contiguous(), d, pad


class _NormScaleFn(torch.autograd.Function):
    """y = scale * x / max(||x||_2, eps) along the last dim (scale may be None)."""

    @staticmethod
    def forward(ctx, x, scale_t, eps):
        K = _KERNELS
        x2, d, pad = _as_rows(x.detach())
        y2 = torch.empty_like(x2)
        inv = torch.empty(x2.shape[0], dtype=torch.float32, device=x.device)
        sdev = None if scale_t is None else scale_t.detach().to(device=x.device, dtype=torch.float32).reshape(1)
        K.l2norm_scale_fwd(x2, y2, inv, sdev, eps)
        ctx.saved = (x2, inv, sdev, d, pad, x.shape, x.dtype, eps)
        y = y2[:, :d] if pad else y2
        return y.reshape(x.shape).to(x.dtype)

    @staticmethod
    def backward(ctx, gy):
        K = _KERNELS
        x2, inv, sdev, d, pad, shape, dtype, eps = ctx.saved
        g2, _, _ = _as_rows(gy.to(x2.dtype) if gy.dtype != x2.dtype else gy)
        gx2 = torch.empty_like(x2)
        dsp = torch.empty(x2.shape[0], dtype=torch.float32, device=x2.device) if sdev is not None else None
        K.l2norm_scale_bwd(x2, g2, inv, gx2, dsp, sdev, eps)
        gx = (gx2[:, :d] if pad else gx2).reshape(shape).to(dtype)
        gs = None
        if sdev is not None and ctx.needs_input_grad[1]:
            t = torch.empty(1, dtype=torch.float32, device=x2.device)
            K.sum_f32(dsp, t)
            gs = t.reshape(())
        return gx, gs, None


class _ScaleFn(torch.autograd.Function):
    """y = scale * x (scale: 0-dim tensor)."""

    @staticmethod
    def forward(ctx, x, scale_t):
        K = _KERNELS
        x2, d, pad = _as_rows(x.detach())
        y2 = torch.empty_like(x2)
        sdev = scale_t.detach().to(device=x.device, dtype=torch.float32).reshape(1)
        K.scale_rows(x2, y2, sdev)
        ctx.saved = (x2, sdev, d, pad, x.shape, x.dtype)
        return (y2[:, :d] if pad else y2).reshape(x.shape).to(x.dtype)

    @staticmethod
    def backward(ctx, gy):
        K = _KERNELS
        x2, sdev, d, pad, shape, dtype = ctx.saved
        g2, _, _ = _as_rows(gy.to(x2.dtype) if gy.dtype != x2.dtype else gy)
        gx2 = torch.empty_like(g2)
        K.scale_rows(g2, gx2, sdev)
        gs = None
        if ctx.needs_input_grad[1]:
            part = torch.empty(x2.shape[0], dtype=torch.float32, device=x2.device)
            K.rowdot(x2, g2, part)
            t = torch.empty(1, dtype=torch.float32, device=x2.device)
            K.sum_f32(part, t)
            gs = t.reshape(())
        return (gx2[:, :d] if pad else gx2).reshape(shape).to(dtype), gs


class Normalize(nn.Module):
    """``F.normalize(x, dim=self.dim, p=2)`` (base_encoder.py:6-12)."""

    def __init__(self, dim: int) -> None:
        super().__init__()
        self.dim = dim

    def forward(self, x):
        last = x.dim() - 1
        dim = self.dim if self.dim >= 0 else x.dim() + self.dim
        if dim != last:
            return _NormScaleFn.apply(x.transpose(dim, last), None, 1e-12).transpose(dim, last)
        return _NormScaleFn.apply(x, None, 1e-12)


class LearnableLogitScaling(nn.Module):
    """``clip(exp(log_logit_scale), max=max_logit_scale) * x`` (base_encoder.py:15-38); same
    constructor, parameter/buffer name and ``extra_repr`` as the reference so checkpoints load."""

    def __init__(self, logit_scale_init: float = 1 / 0.07, learnable: bool = True, max_logit_scale: float = 100) -> None:
        super().__init__()
        self.max_logit_scale = max_logit_scale
        self.logit_scale_init = logit_scale_init
        self.learnable = learnable
        log_logit_scale = torch.ones([]) * np.log(self.logit_scale_init)
        if learnable:
            self.log_logit_scale = nn.Parameter(log_logit_scale)
        else:
            self.register_buffer("log_logit_scale", log_logit_scale)

    def effective_scale(self):
        return torch.clip(self.log_logit_scale.exp(), max=self.max_logit_scale)

    def forward(self, x):
        return _ScaleFn.apply(x, self.effective_scale())

    def extra_repr(self):
        st = f"logit_scale_init={self.logit_scale_init},learnable={self.learnable}," \
             f" max_logit_scale={self.max_logit_scale}"
        return st


class NormalizeAndScale(nn.Module):
    """Fused ``Normalize(dim=-1)`` -> ``LearnableLogitScaling`` (the ``norm`` Sequential the
    reference builds at base_encoder.py:171-178): one read and one write of the embedding."""

    def __init__(self, logit_scale_init: float = 1 / 0.07, learnable: bool = True, max_logit_scale: float = 100) -> None:
        super().__init__()
        self.scaling = LearnableLogitScaling(logit_scale_init, learnable, max_logit_scale)

    def forward(self, x):
        return _NormScaleFn.apply(x, self.scaling.effective_scale(), 1e-12)
