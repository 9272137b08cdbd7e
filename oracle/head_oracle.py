"""CPU oracle for the projection heads in front of the ClipLoss path.  TEST INFRASTRUCTURE ONLY
(same rules as oracle/clip_oracle.py: only tests/, smoke() and bench.py's CPU leg may import it).

Numpy float64 restatement, forward AND backward, of the reference's ``BaseEncoder`` head
(/root/reference/src/models/components/base_encoder.py:107-194):

    pooling (MeanPooling :107-118 | CLSTokenPooling :121-126 | Attention1dPooling :84-104 | identity)
    -> proj ('linear': LayerNorm, Linear(no bias) :146-150 | 'mlp': LayerNorm, Linear, GELU, LayerNorm, Linear :151-159)
    -> norm (F.normalize(dim=-1) :6-12 [-> clip(exp(log_s), max) * x :15-33])

Parity status: PINNED by tests/golden/head_*.npz, produced by importing the unmodified reference
``BaseEncoder`` (oracle/make_golden.py::head_cases) - tests/test_oracle_golden.py checks this file
against them.  torch.nn.LayerNorm / GELU / Linear are third-party arithmetic (torch); their
published definitions are restated here: LayerNorm = biased variance with eps inside the root,
GELU = x Phi(x) with the exact erf.
"""
from __future__ import annotations

import math

import numpy as np

_erf = np.vectorize(math.erf, otypes=[np.float64])


def meanpool_fwd(x, mask):
    """x: (B, L, D); mask: (B, L) or None -> (B, D)   (base_encoder.py:111-118)"""
    if x.ndim == 2:
        return x
    if mask is None:
        return x.mean(axis=1)
    m = mask.astype(np.float64)
    return (x * m[:, :, None]).sum(axis=1) / m.sum(axis=1, keepdims=True)


def meanpool_bwd(gy, mask, shape):
    B, L, D = shape
    if mask is None:
        return np.broadcast_to(gy[:, None, :] / L, shape).copy()
    m = mask.astype(np.float64)
    return gy[:, None, :] * (m / m.sum(axis=1, keepdims=True))[:, :, None]


def attnpool_fwd(x, mask, w, bias):
    """Attention1dPooling.forward (base_encoder.py:89-104): x (B, L, D), w (D,), bias scalar."""
    s = x @ w + bias
    if mask is not None:
        s = np.where(mask.astype(bool), s, -np.inf)
    s = s - s.max(axis=1, keepdims=True)
    p = np.exp(s)
    p = p / p.sum(axis=1, keepdims=True)
    return (p[:, :, None] * x).sum(axis=1), p


def attnpool_bwd(gy, x, w, p):
    dp = (x * gy[:, None, :]).sum(axis=2)
    ds = p * (dp - (p * dp).sum(axis=1, keepdims=True))
    gx = p[:, :, None] * gy[:, None, :] + ds[:, :, None] * w[None, None, :]
    return gx, (ds[:, :, None] * x).sum(axis=(0, 1)), ds.sum()


def layernorm_fwd(x, gamma, beta, eps=1e-5):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    xhat = (x - mu) * rstd
    return xhat * gamma + beta, (xhat, rstd)


def layernorm_bwd(gy, gamma, cache):
    xhat, rstd = cache
    g = gy * gamma
    gx = rstd * (g - g.mean(axis=-1, keepdims=True) - xhat * (g * xhat).mean(axis=-1, keepdims=True))
    return gx, (gy * xhat).reshape(-1, gy.shape[-1]).sum(axis=0), gy.reshape(-1, gy.shape[-1]).sum(axis=0)


def gelu_fwd(x):
    return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))


def gelu_bwd(gy, x):
    cdf = 0.5 * (1.0 + _erf(x / math.sqrt(2.0)))
    pdf = np.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)
    return gy * (cdf + x * pdf)


def normalize_fwd(x, eps=1e-12):
    nrm = np.maximum(np.sqrt((x * x).sum(axis=-1, keepdims=True)), eps)
    return x / nrm, nrm


def normalize_bwd(gy, x, nrm, eps=1e-12):
    y = x / nrm
    proj = np.where(nrm > eps, (y * gy).sum(axis=-1, keepdims=True), 0.0)
    return (gy - y * proj) / nrm


def head_forward_backward(x, mask, params, *, proj_type, pooling_type, use_logit_scale, gy, max_logit_scale=100.0):
    """Value and gradients of the BaseEncoder head (base_encoder.py:190-194) in float64.

    params: dict with the reference's state_dict names ('proj.0.weight', 'proj.0.bias', 'proj.1.weight',
    ['proj.3.weight', 'proj.3.bias', 'proj.4.weight'], ['norm.1.log_logit_scale']).
    Returns (y, grads) with grads['x'] and one entry per parameter name."""
    P = {k: np.asarray(v, dtype=np.float64) for k, v in params.items()}
    x = np.asarray(x, dtype=np.float64)
    gy = np.asarray(gy, dtype=np.float64)
    # ---- forward
    if pooling_type == "mean":
        h0 = meanpool_fwd(x, mask)
    elif pooling_type == "cls":
        h0 = x[:, 0]
    elif pooling_type == "attention1d":
        aw = P["pooling.layer.weight"].reshape(-1)
        h0, ap = attnpool_fwd(x, mask, aw, float(P["pooling.layer.bias"].reshape(-1)[0]))
    else:
        h0 = x
    grads = {}
    if proj_type in ("linear", "mlp"):
        a1, c1 = layernorm_fwd(h0, P["proj.0.weight"], P["proj.0.bias"])
        z1 = a1 @ P["proj.1.weight"].T
        if proj_type == "mlp":
            a2 = gelu_fwd(z1)
            a3, c3 = layernorm_fwd(a2, P["proj.3.weight"], P["proj.3.bias"])
            z = a3 @ P["proj.4.weight"].T
        else:
            z = z1
    else:
        z = h0
    yn, nrm = normalize_fwd(z)
    if use_logit_scale:
        s = min(math.exp(float(P["norm.1.log_logit_scale"])), max_logit_scale)
        y = s * yn
    else:
        s, y = 1.0, yn
    # ---- backward
    if use_logit_scale:
        # d/d log_s of clip(exp(log_s), max): exp(log_s) below the clip, 0 above (torch.clip's subgradient)
        ds = (gy * yn).sum()
        grads["norm.1.log_logit_scale"] = ds * (s if math.exp(float(P["norm.1.log_logit_scale"])) <= max_logit_scale else 0.0)
    gz = normalize_bwd(s * gy, z, nrm)
    if proj_type in ("linear", "mlp"):
        if proj_type == "mlp":
            grads["proj.4.weight"] = gz.T @ a3
            ga3 = gz @ P["proj.4.weight"]
            ga2, grads["proj.3.weight"], grads["proj.3.bias"] = layernorm_bwd(ga3, P["proj.3.weight"], c3)
            gz1 = gelu_bwd(ga2, z1)
        else:
            gz1 = gz
        grads["proj.1.weight"] = gz1.T @ a1
        ga1 = gz1 @ P["proj.1.weight"]
        gh0, grads["proj.0.weight"], grads["proj.0.bias"] = layernorm_bwd(ga1, P["proj.0.weight"], c1)
    else:
        gh0 = gz
    if pooling_type == "mean" and x.ndim == 3:
        grads["x"] = meanpool_bwd(gh0, mask, x.shape)
    elif pooling_type == "cls":
        gx = np.zeros_like(x)
        gx[:, 0] = gh0
        grads["x"] = gx
    elif pooling_type == "attention1d":
        grads["x"], gw, gb = attnpool_bwd(gh0, x, aw, ap)
        grads["pooling.layer.weight"] = gw.reshape(P["pooling.layer.weight"].shape)
        grads["pooling.layer.bias"] = np.asarray(gb).reshape(P["pooling.layer.bias"].shape)
    else:
        grads["x"] = gh0
    return y, grads
