"""Projection heads in front of the ClipLoss path on B200.

Drop-in for the reference's ``BaseEncoder`` head (``src/models/components/base_encoder.py:107-194``):
``pooling -> proj -> norm`` with the same constructor, the same sub-module names and the same
parameter names, so that a reference checkpoint's ``pooling.* / proj.* / norm.*`` keys load:

    pooling   MeanPooling (masked mean, :107-118) | CLSTokenPooling (:121-126) | Attention1dPooling (:84-104) | Identity
    proj      Identity | LayerNorm -> Linear(no bias)                                ('linear', :146-150)
                       | LayerNorm -> Linear -> GELU -> LayerNorm -> Linear         ('mlp',    :151-159)
    norm      Normalize(dim=-1) [-> LearnableLogitScaling]                           (:166-176, epilogue.py)

It is the only trainable part of OneProt when the language-model towers are frozen
(``configs/model/components/sequence.yaml:12``), i.e. what the loss gradient flows into.

What runs where: pooling, LayerNorm and GELU are HBM-bound row kernels of ``liboneprot_clip.so``
(``csrc/head_kernels.cu``), forward and backward; the Linear layers are the tcgen05 GEMM of the loss
path (``oneprot_gemm_bf16_ex``: y = x W^T, dx = gy W, dW = gy^T x).  bf16 modules run the GEMMs on
bf16 operands with fp32 accumulation; fp32 modules (the reference's default precision, TF32 matmuls,
``src/train.py:98``) split both operands into bf16 limbs (hh + hm + mh, error 2^-16 per product -
tighter than TF32's 2^-11) exactly as ``ClipLoss`` does for fp32 features.  No eager fallback.

``Attention1dPooling`` (:84-104, the sequence tower of ``configs/experiment/train_ddp_1.yaml``): one dot
product per token (warp per token), a masked softmax over the tokens (block per batch row) and the
weighted sum through the pooling kernel; its backward reuses the same three kernels.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import kernels as _cuda_kernels
from .epilogue import LearnableLogitScaling, Normalize, _NormScaleFn

_KERNELS = _cuda_kernels

_DTYPES = (torch.bfloat16, torch.float32)


def _rows(x: torch.Tensor):
    """(rows, d) contiguous view of x over its last dim."""
    d = x.shape[-1]
    return x.reshape(-1, d).contiguous(), d


def _check_dtype(x, what):
    if x.dtype not in _DTYPES:
        raise ValueError(f"{what}: bf16 or fp32 expected, got {x.dtype}")


# ---------------------------------------------------------------------------------------------
# LayerNorm
# ---------------------------------------------------------------------------------------------
class _LayerNormFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        K = _KERNELS
        x2, d = _rows(x.detach())
        w = weight.detach().to(x2.dtype).contiguous()
        b = bias.detach().to(x2.dtype).contiguous()
        y2 = torch.empty_like(x2)
        mean = torch.empty(x2.shape[0], dtype=torch.float32, device=x.device)
        rstd = torch.empty(x2.shape[0], dtype=torch.float32, device=x.device)
        K.layernorm_fwd(x2, w, b, y2, mean, rstd, eps)
        ctx.saved = (x2, w, mean, rstd, x.shape, weight.dtype, bias.dtype)
        return y2.reshape(x.shape)

    @staticmethod
    def backward(ctx, gy):
        K = _KERNELS
        x2, w, mean, rstd, shape, wdt, bdt = ctx.saved
        g2, d = _rows(gy.to(x2.dtype) if gy.dtype != x2.dtype else gy)
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        gx2 = torch.empty_like(x2) if need_x else None
        dg = db = None
        if need_w or need_b:
            dg = torch.empty(d, dtype=torch.float32, device=x2.device)
            db = torch.empty(d, dtype=torch.float32, device=x2.device)
        K.layernorm_bwd(x2, g2, w, mean, rstd, gx2, dg, db)
        return (gx2.reshape(shape) if need_x else None, dg.to(wdt) if need_w else None, db.to(bdt) if need_b else None, None)


class LayerNorm(nn.Module):
    """``torch.nn.LayerNorm(d)`` over the last dim with affine parameters ``weight`` / ``bias``."""

    def __init__(self, normalized_shape: int, eps: float = 1e-5):
        super().__init__()
        if not isinstance(normalized_shape, int):
            (normalized_shape,) = tuple(normalized_shape)
        if normalized_shape % 8:
            raise ValueError("oneprot_b200.LayerNorm: the normalised dim must be a multiple of 8")
        self.normalized_shape = (normalized_shape,)
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))

    def forward(self, x):
        _check_dtype(x, "LayerNorm")
        return _LayerNormFn.apply(x, self.weight, self.bias, self.eps)

    def extra_repr(self):
        return f"{self.normalized_shape}, eps={self.eps}, elementwise_affine=True"


# ---------------------------------------------------------------------------------------------
# Linear (no bias) on the tcgen05 GEMM
# ---------------------------------------------------------------------------------------------
def _limbs(x2: torch.Tensor, side: int):
    """fp32 (rows, d) -> bf16 (rows, 3 d): side 0 = [h | h | m], side 1 = [h | m | h]."""
    out = torch.empty(x2.shape[0], 3 * x2.shape[1], dtype=torch.bfloat16, device=x2.device)
    _KERNELS.split_fp32(x2, out, side, 3)
    return out


def _gemm_terms(terms, M, Nc, K, out):
    """out = sum of op(A) op(B) over `terms` [(A, a_mn, B, b_mn)], chained through the fp32 accumulator."""
    Kn = _KERNELS
    f32 = out.dtype == torch.float32
    acc = out if f32 else (torch.empty(M, Nc, dtype=torch.float32, device=out.device) if len(terms) > 1 else None)
    for i, (A, a_mn, B, b_mn) in enumerate(terms):
        first, last = i == 0, i == len(terms) - 1
        kw = dict(acc_in=None if first else acc)
        if last and not f32:
            Kn.gemm_bf16(A, a_mn, B, b_mn, M, Nc, K, out=out, **kw)
        else:
            Kn.gemm_bf16(A, a_mn, B, b_mn, M, Nc, K, acc_out=acc, **kw)


class _LinearFn(torch.autograd.Function):
    """y = x W^T for x (rows, in), W (out, in); no bias."""

    @staticmethod
    def forward(ctx, x, weight):
        x2, d_in = _rows(x.detach())
        W = weight.detach()
        if W.dtype != x2.dtype:
            W = W.to(x2.dtype)
        W = W.contiguous()
        n, d_out = x2.shape[0], W.shape[0]
        y2 = torch.empty(n, d_out, dtype=x2.dtype, device=x.device)
        if x2.dtype == torch.bfloat16:
            _gemm_terms([(x2, False, W, False)], n, d_out, d_in, y2)
            ctx.saved = (x2, W, None, None)
        else:
            xs, ws = _limbs(x2, 0), _limbs(W, 1)          # one GEMM over K = 3 in: hh + hm + mh
            _gemm_terms([(xs, False, ws, False)], n, d_out, 3 * d_in, y2)
            ctx.saved = (x2, W, xs, ws)
        ctx.meta = (x.shape, weight.dtype)
        return y2.reshape(x.shape[:-1] + (d_out,))

    @staticmethod
    def backward(ctx, gy):
        x2, W, xs, ws = ctx.saved
        shape, wdt = ctx.meta
        n, d_in = x2.shape
        d_out = W.shape[0]
        g2, _ = _rows(gy.to(x2.dtype) if gy.dtype != x2.dtype else gy)
        gx = gw = None
        if x2.dtype == torch.bfloat16:
            if ctx.needs_input_grad[0]:
                gx = torch.empty(n, d_in, dtype=torch.bfloat16, device=x2.device)
                _gemm_terms([(g2, False, W, True)], n, d_in, d_out, gx)            # gy (n x out) . W (out x in)
            if ctx.needs_input_grad[1]:
                gw = torch.empty(d_out, d_in, dtype=torch.bfloat16, device=x2.device)
                _gemm_terms([(g2, True, x2, True)], d_out, d_in, n, gw)            # gy^T (out x n) . x (n x in)
        else:
            gs = _limbs(g2, 0)                                                     # [h | h | m] along `out`
            g_h, g_m = gs[:, 0:d_out], gs[:, 2 * d_out:3 * d_out]
            if ctx.needs_input_grad[0]:
                w_h, w_m = ws[:, 0:d_in], ws[:, d_in:2 * d_in]                      # [h | m | h] along `in`
                gx = torch.empty(n, d_in, dtype=torch.float32, device=x2.device)
                _gemm_terms([(g_h, False, w_h, True), (g_h, False, w_m, True), (g_m, False, w_h, True)], n, d_in, d_out, gx)
            if ctx.needs_input_grad[1]:
                x_h, x_m = xs[:, 0:d_in], xs[:, 2 * d_in:3 * d_in]
                gw = torch.empty(d_out, d_in, dtype=torch.float32, device=x2.device)
                _gemm_terms([(g_h, True, x_h, True), (g_h, True, x_m, True), (g_m, True, x_h, True)], d_out, d_in, n, gw)
        if gx is not None:
            gx = gx.reshape(shape)
        if gw is not None and gw.dtype != wdt:
            gw = gw.to(wdt)
        return gx, gw


class Linear(nn.Module):
    """``torch.nn.Linear(in_features, out_features, bias=False)`` on the tcgen05 GEMM."""

    def __init__(self, in_features: int, out_features: int, bias: bool = False):
        super().__init__()
        if bias:
            raise NotImplementedError("oneprot_b200.Linear: the reference's projection heads have no bias (base_encoder.py:149-158)")
        if in_features % 8 or out_features % 8:
            raise ValueError("oneprot_b200.Linear: in_features and out_features must be multiples of 8 (16-byte TMA row pitch)")
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.register_parameter("bias", None)
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))     # torch.nn.Linear.reset_parameters

    def forward(self, x):
        _check_dtype(x, "Linear")
        return _LinearFn.apply(x, self.weight)

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, bias=False"


# ---------------------------------------------------------------------------------------------
# GELU
# ---------------------------------------------------------------------------------------------
class _GeluFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x):
        x2 = x.detach().contiguous()
        if x2.numel() % 8:
            raise ValueError("oneprot_b200.GELU: element count must be a multiple of 8")
        y = torch.empty_like(x2)
        _KERNELS.gelu(x2, y)
        ctx.saved = x2
        return y

    @staticmethod
    def backward(ctx, gy):
        x2 = ctx.saved
        g = (gy.to(x2.dtype) if gy.dtype != x2.dtype else gy).contiguous()
        gx = torch.empty_like(x2)
        _KERNELS.gelu(x2, gx, g)
        return gx


class GELU(nn.Module):
    """Exact (erf) GELU, ``torch.nn.GELU()``."""

    def forward(self, x):
        _check_dtype(x, "GELU")
        return _GeluFn.apply(x)


# ---------------------------------------------------------------------------------------------
# pooling
# ---------------------------------------------------------------------------------------------
class _MeanPoolFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, mask):
        K = _KERNELS
        x3 = x.detach().contiguous()
        B, L, D = x3.shape
        if D % 8:
            raise ValueError("oneprot_b200.MeanPooling: the feature dim must be a multiple of 8")
        m = None if mask is None else mask.detach().to(device=x.device, dtype=torch.float32).reshape(B, L).contiguous()
        y = torch.empty(B, D, dtype=x3.dtype, device=x.device)
        inv = torch.empty(B, dtype=torch.float32, device=x.device)
        K.meanpool_fwd(x3, m, y, inv)
        ctx.saved = (m, inv, x3.shape)
        return y

    @staticmethod
    def backward(ctx, gy):
        m, inv, shape = ctx.saved
        g = gy.contiguous()
        gx = torch.empty(shape, dtype=g.dtype, device=g.device)
        _KERNELS.meanpool_bwd(g, m, inv, gx)
        return gx, None


class MeanPooling(nn.Module):
    """base_encoder.py:107-118: 2-D features pass through; masked mean over dim 1 otherwise."""

    def forward(self, features, input_mask=None):
        if features.dim() == 2:
            return features
        _check_dtype(features, "MeanPooling")
        return _MeanPoolFn.apply(features, input_mask)


class _AttnPoolFn(torch.autograd.Function):
    """out[b] = sum_l softmax_l(<w, x[b,l]> + bias, masked)[l] * x[b,l]   (base_encoder.py:84-104)"""

    @staticmethod
    def forward(ctx, x, weight, bias, mask):
        K = _KERNELS
        x3 = x.detach().contiguous()
        B, L, D = x3.shape
        if D % 8:
            raise ValueError("oneprot_b200.Attention1dPooling: the feature dim must be a multiple of 8")
        w = weight.detach().reshape(D).to(x3.dtype).contiguous()
        bs = bias.detach().reshape(1).to(device=x.device, dtype=torch.float32)
        m = None if mask is None else mask.detach().to(device=x.device, dtype=torch.float32).reshape(B, L).contiguous()
        p = torch.empty(B, L, dtype=torch.float32, device=x.device)
        K.token_dot(x3, w, p, bias=bs, mask=m)
        K.softmax_rows(p, p)
        y = torch.empty(B, D, dtype=x3.dtype, device=x.device)
        K.meanpool_fwd(x3, p, y, None, normalize=False)
        ctx.saved = (x3, w, p)
        ctx.meta = (weight.shape, weight.dtype, bias.shape, bias.dtype)
        return y

    @staticmethod
    def backward(ctx, gy):
        K = _KERNELS
        x3, w, p = ctx.saved
        wshape, wdt, bshape, bdt = ctx.meta
        B, L, D = x3.shape
        g = (gy.to(x3.dtype) if gy.dtype != x3.dtype else gy).contiguous()
        dp = torch.empty(B, L, dtype=torch.float32, device=g.device)
        K.token_dot(x3, g, dp)                         # d p[b,l] = <g[b], x[b,l]>
        ds = torch.empty_like(dp)
        K.softmax_rows_bwd(p, dp, ds)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x3)
            K.attnpool_bwd_x(g, p, ds, w, gx)          # p[b,l] g[b] + ds[b,l] w
        if ctx.needs_input_grad[1]:
            # d w = sum_{b,l} ds[b,l] x[b,l]: per-batch weighted sums, then a fixed-order sum over the batch
            part = torch.empty(B, D, dtype=x3.dtype, device=g.device)
            K.meanpool_fwd(x3, ds, part, None, normalize=False)
            gwv = torch.empty(D, dtype=torch.float32, device=g.device)
            K.sum_slots_f32(part if part.dtype == torch.float32 else part.float(), gwv)
            gw = gwv.reshape(wshape).to(wdt)
        if ctx.needs_input_grad[2]:
            t = torch.empty(1, dtype=torch.float32, device=g.device)
            K.sum_f32(ds.reshape(-1), t)
            gb = t.reshape(bshape).to(bdt)
        return gx, gw, gb, None


class _ScoreLayer(nn.Module):
    """Parameters of the reference's ``MaskedConv1d(hidden_size, 1, 1)`` (base_encoder.py:40-81,87): a
    1x1 convolution to one channel = one dot product per token.  Same names, shapes and initialisation
    as ``torch.nn.Conv1d`` so that ``pooling.layer.weight / bias`` of a reference checkpoint load."""

    def __init__(self, hidden_size: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(1, hidden_size, 1))
        self.bias = nn.Parameter(torch.empty(1))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(hidden_size)
        nn.init.uniform_(self.bias, -bound, bound)


class Attention1dPooling(nn.Module):
    """base_encoder.py:84-104: learned scores, masked softmax over the tokens, weighted sum."""

    def __init__(self, hidden_size: int):
        super().__init__()
        self.layer = _ScoreLayer(hidden_size)

    def forward(self, x, input_mask=None):
        _check_dtype(x, "Attention1dPooling")
        return _AttnPoolFn.apply(x, self.layer.weight, self.layer.bias, input_mask)


class CLSTokenPooling(nn.Module):
    """base_encoder.py:121-126 (a strided view; no arithmetic)."""

    def forward(self, features, input_mask=None):
        return features[:, 0]


class _IdentityPooling(nn.Identity):
    """``nn.Identity`` that tolerates the mask argument some encoders pass (msa_encoder.py:49)."""

    def forward(self, features, input_mask=None):
        return features


class _NormSequential(nn.Sequential):
    """The reference's ``norm`` Sequential (base_encoder.py:171-178) with the same children and
    state_dict keys (``norm.1.log_logit_scale``).  [Normalize(dim=-1), LearnableLogitScaling] run as ONE
    fused kernel (one read and one write of the embedding instead of two of each); anything else runs
    child by child.  Encoders that call ``self.norm(projected)`` themselves (sequence_encoder.py:81) get
    the fusion as well."""

    def forward(self, x):
        if (len(self) == 2 and isinstance(self[0], Normalize) and isinstance(self[1], LearnableLogitScaling)
                and self[0].dim in (-1, x.dim() - 1)):
            return _NormScaleFn.apply(x, self[1].effective_scale(), 1e-12)
        return super().forward(x)


# ---------------------------------------------------------------------------------------------
# BaseEncoder head
# ---------------------------------------------------------------------------------------------
class BaseEncoder(nn.Module):
    """Same constructor, attributes and ``forward(x, input_mask=None)`` as the reference
    (base_encoder.py:129-194).  Sub-classes add the tower and call ``self.pooling / self.proj /
    self.norm`` exactly like the reference encoders do (sequence_encoder.py:76-81)."""

    def __init__(self, d_model: int, output_dim: int, proj_type: str = None, use_logit_scale: bool = False,
                 learnable_logit_scale: bool = False, pooling_type: str = 'mean'):
        super().__init__()
        self.d_model = d_model
        self.output_dim = output_dim
        self.pooling_type = pooling_type
        self.proj = self._create_projection(proj_type)
        self.norm = self._create_normalization(use_logit_scale, learnable_logit_scale)
        self.pooling = self._create_pooling(pooling_type)

    def _create_projection(self, proj_type):
        if (self.d_model == self.output_dim) and (proj_type is None):
            return nn.Sequential(nn.Identity())
        elif proj_type == 'linear':
            return nn.Sequential(LayerNorm(self.d_model), Linear(self.d_model, self.output_dim, bias=False))
        elif proj_type == 'mlp':
            hidden_size = (self.d_model + self.output_dim) // 2
            return nn.Sequential(LayerNorm(self.d_model), Linear(self.d_model, hidden_size, bias=False), GELU(),
                                 LayerNorm(hidden_size), Linear(hidden_size, self.output_dim, bias=False))
        else:
            return nn.Sequential(nn.Identity())

    def _create_normalization(self, use_logit_scale, learnable_logit_scale=False):
        layers = [Normalize(dim=-1)]
        if use_logit_scale:
            layers.append(LearnableLogitScaling(learnable=bool(learnable_logit_scale)))
        return _NormSequential(*layers)

    def _create_pooling(self, pooling_type, hidden_size=1280):
        if pooling_type == 'mean':
            return MeanPooling()
        elif pooling_type == 'cls':
            return CLSTokenPooling()
        elif pooling_type == 'attention1d':
            return Attention1dPooling(hidden_size)
        else:
            return _IdentityPooling()

    def forward(self, x, input_mask=None):
        x = self.pooling(x, input_mask)
        x = self.proj(x)
        x = self.norm(x)
        return x
