"""CPU oracle for the OneProt ClipLoss hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  The product package ``oneprot_b200`` never does:
it fails loudly when its CUDA extension is missing.

Parity status: PINNED.  Every function here is checked (tests/test_oracle_golden.py) against
fixtures under ``tests/golden/`` that were produced by importing the *unmodified* reference
(``/root/reference/src/models/components/loss.py`` and ``base_encoder.py``) in the build
container, single-process and 2-rank gloo, by ``oracle/make_golden.py`` (committed).  The
reference's own test-suite holds no golden vector for this path, so those
reference-generated fixtures are the pin.

Two restatements are kept on purpose:

* ``clip_loss_port`` - a torch (CPU) port that performs the same sequence of library ops the
  reference performs (scale-then-matmul, ``cross_entropy`` in both directions, autograd
  backward).  This is what ``bench.py`` times as the CPU baseline (kind "port"): it uses the
  host cores exactly the way the reference does (oneDNN matmul + ATen softmax).
* ``clip_loss_closed_form`` - an independent numpy float64 closed form of value and gradients
  for every (local_loss, gather_with_grad, world_size) convention of the reference, so the
  CUDA path is never checked against a re-run of its own formulas.

Reference citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np


# ----------------------------------------------------------------------------------------------
# numpy float64 closed form
# ----------------------------------------------------------------------------------------------
def _lse(z: np.ndarray, axis: int) -> np.ndarray:
    m = z.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(z - m).sum(axis=axis, keepdims=True))).squeeze(axis)


@dataclass
class ClipResult:
    loss: float          # value returned on this rank (src/models/components/loss.py:109-114)
    dA: np.ndarray       # gradient w.r.t. this rank's first positional feature tensor  (n x d)
    dB: np.ndarray       # gradient w.r.t. this rank's second positional feature tensor (n x d)
    dscale: float        # gradient w.r.t. logit_scale seen on this rank
    row_lse: np.ndarray  # log-sum-exp over columns, natural log, for the rows this rank scores
    col_lse: np.ndarray  # log-sum-exp over rows for the columns this rank scores


def clip_loss_closed_form(A_all: np.ndarray, B_all: np.ndarray, scale: float, *, rank: int = 0,
                          world_size: int = 1, local_loss: bool = False,
                          gather_with_grad: bool = False,
                          grad_outputs: Optional[np.ndarray] = None) -> ClipResult:
    """Value and gradients the reference produces on ``rank`` (float64).

    ``A_all``/``B_all`` are the concatenation over ranks of the per-rank feature tensors, in
    rank order - exactly what ``gather_features`` builds (loss.py:19-46).  Z = (scale*A) B^T
    (loss.py:92-99; with float64 the scale-then-round of the reference is exact).

    Conventions (measured on the unmodified reference):
      * local_loss=False: value = global loss L_g on every rank (loss.py:95-96,109-112).
        gather_with_grad=False: dA_r = g_r * dL_g/dA_r  (only the local slot carries grad,
        loss.py:39-42).  gather_with_grad=True: dA_r = sum_r' g_r' dL_g/dA_r - the backward of
        ``torch.distributed.nn.all_gather`` is a reduce-scatter SUM (loss.py:32-33).
      * local_loss=True: value = L_r, the mean over this rank's rows of logits_per_modality
        and of logits_per_sequence (loss.py:92-93 with labels offset by n*rank, loss.py:76-77).
        gather_with_grad=True: dA_r = sum_r' g_r' dL_r'/dA_r.  gather_with_grad=False: only the
        query side carries grad (the gathered copies are constants, loss.py:35-38).
      * d logit_scale on rank r = g_r * d(value_r)/d scale.
    ``grad_outputs[r']`` is the upstream gradient of rank r' (default: all ones, i.e.
    ``loss.backward()`` on every rank).
    """
    A = np.asarray(A_all, dtype=np.float64)
    B = np.asarray(B_all, dtype=np.float64)
    N, d = A.shape
    W = int(world_size)
    assert B.shape == (N, d) and N % W == 0
    n = N // W
    g = np.ones(W) if grad_outputs is None else np.asarray(grad_outputs, dtype=np.float64)
    s = float(scale)

    dot = A @ B.T                      # a_i . b_j
    Z = s * dot
    rl = _lse(Z, 1)                    # row LSE  (logits_per_modality rows)
    cl = _lse(Z, 0)                    # col LSE  (logits_per_sequence rows)
    diag = np.diag(Z)
    P = np.exp(Z - rl[:, None])        # row softmax
    Q = np.exp(Z - cl[None, :])        # column softmax
    I = np.eye(N)
    R = slice(rank * n, (rank + 1) * n)
    owner = np.repeat(np.arange(W), n)  # rank that owns row/col index

    if W == 1 or not local_loss:
        loss = 0.5 * ((rl - diag).mean() + (cl - diag).mean())
        dZ_unit = (P + Q - 2 * I) / (2 * N)          # dL_g/dZ
        if W == 1 or not gather_with_grad:
            gz = g[rank] * dZ_unit
        else:
            gz = g.sum() * dZ_unit
        dA = s * (gz @ B)[R]
        dB = s * (gz.T @ A)[R]
        dscale = g[rank] * float((dZ_unit * dot).sum())
        row_lse, col_lse = rl, cl
    else:
        loss = 0.5 * ((rl[R] - diag[R]).mean() + (cl[R] - diag[R]).mean())
        # dL_r'/dZ: rows of rank r' carry (P - I)/(2n), columns of rank r' carry (Q - I)/(2n)
        if gather_with_grad:
            gz = (g[owner][:, None] * (P - I) + g[owner][None, :] * (Q - I)) / (2 * n)
            dA = s * (gz @ B)[R]
            dB = s * (gz.T @ A)[R]
        else:
            gzA = np.zeros_like(Z)
            gzA[R, :] = g[rank] * (P - I)[R, :] / (2 * n)     # grad only through A_r (query)
            gzB = np.zeros_like(Z)
            gzB[:, R] = g[rank] * (Q - I)[:, R] / (2 * n)     # grad only through B_r (query)
            dA = s * (gzA @ B)[R]
            dB = s * (gzB.T @ A)[R]
        own = np.zeros_like(Z)
        own[R, :] += (P - I)[R, :] / (2 * n)
        own[:, R] += (Q - I)[:, R] / (2 * n)
        dscale = g[rank] * float((own * dot).sum())
        row_lse, col_lse = rl[R], cl[R]
    return ClipResult(float(loss), dA, dB, float(dscale), row_lse, col_lse)


def normalize_closed_form(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """``Normalize.forward`` = F.normalize(x, dim=-1, p=2) (base_encoder.py:6-12)."""
    x = np.asarray(x, dtype=np.float64)
    nrm = np.maximum(np.sqrt((x * x).sum(-1, keepdims=True)), eps)
    return x / nrm


def normalize_backward_closed_form(x: np.ndarray, gy: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """Backward of F.normalize: (gy - y (y.gy)) / max(||x||, eps) (clamped branch: gy/eps)."""
    x = np.asarray(x, dtype=np.float64)
    gy = np.asarray(gy, dtype=np.float64)
    nrm = np.sqrt((x * x).sum(-1, keepdims=True))
    y = x / np.maximum(nrm, eps)
    gx = (gy - y * (y * gy).sum(-1, keepdims=True)) / np.maximum(nrm, eps)
    return np.where(nrm > eps, gx, gy / eps)


def logit_scaling_closed_form(x: np.ndarray, log_logit_scale: float, max_logit_scale: float = 100.0):
    """``LearnableLogitScaling.forward`` = clip(exp(log_s), max=max) * x (base_encoder.py:15-33).

    Returns (y, effective_scale)."""
    s = min(math.exp(log_logit_scale), max_logit_scale)
    return s * np.asarray(x, dtype=np.float64), s


def retrieval_metric_closed_form(sequence_outputs: np.ndarray, modality_outputs: np.ndarray, ks=(1, 10, 100)) -> dict:
    """``RetrievalMetric.compute`` (src/models/components/retrieval_metric.py:76-102) restated with
    numpy: similarity S M^T, descending argsort, position of the label, floor(median)+1 and R@k."""
    S = np.asarray(sequence_outputs, dtype=np.float64)
    M = np.asarray(modality_outputs, dtype=np.float64)
    out = {}
    logits = {"seq_to_mod": S @ M.T, "mod_to_seq": (S @ M.T).T}
    gt = np.arange(len(M)).reshape(-1, 1)
    for name, z in logits.items():
        ranking = np.argsort(-z, axis=1, kind="stable")
        preds = np.where(ranking == gt)[1]
        out[f"{name}_median_rank"] = np.floor(np.median(preds)) + 1
        for k in ks:
            out[f"{name}_R@{k}"] = np.mean(preds < k)
    return out


# ----------------------------------------------------------------------------------------------
# torch CPU port (same library ops as the reference; the timed CPU baseline)
# ----------------------------------------------------------------------------------------------
def clip_loss_port(A, B, logit_scale=1.0):
    """world_size == 1 port of ``ClipLoss.forward`` (loss.py:85-114): two GEMMs with the scale
    multiplied into the left operand first (loss.py:98-99, Python precedence), labels =
    arange (loss.py:72-83), mean-reduced cross entropy in both directions averaged
    (loss.py:109-112).  Returns a tensor attached to the autograd graph of A, B, logit_scale."""
    import torch
    import torch.nn.functional as F

    z_ab = (logit_scale * A) @ B.T
    z_ba = (logit_scale * B) @ A.T
    target = torch.arange(z_ab.shape[0], device=A.device, dtype=torch.long)
    return (F.cross_entropy(z_ab, target) + F.cross_entropy(z_ba, target)) / 2


def clip_loss_port_panel(A, B, m, logit_scale=1.0):
    """The port restricted to the first m rows of both logit matrices (a bounded, row-separable
    sample of the N x N work for the CPU baseline of bench.py): same ops as ``clip_loss_port``
    (loss.py:98-99,109-112) on z_ab[:m] and z_ba[:m]."""
    import torch
    import torch.nn.functional as F

    z_ab = (logit_scale * A[:m]) @ B.T
    z_ba = (logit_scale * B[:m]) @ A.T
    target = torch.arange(m, device=A.device, dtype=torch.long)
    return (F.cross_entropy(z_ab, target) + F.cross_entropy(z_ba, target)) / 2


def clip_loss_port_fwd_bwd(A, B, logit_scale=1.0):
    """One fwd+bwd step of the port; returns (loss, dA, dB) as detached tensors."""
    A = A.detach().requires_grad_(True)
    B = B.detach().requires_grad_(True)
    loss = clip_loss_port(A, B, logit_scale)
    loss.backward()
    return loss.detach(), A.grad, B.grad


def clip_loss_port_distributed(a_loc, b_loc, logit_scale, *, rank, world_size, local_loss,
                               gather_with_grad):
    """Multi-rank port (needs an initialised torch.distributed group): gather_features
    (loss.py:19-46) + get_logits (loss.py:85-101) + forward (loss.py:103-114)."""
    import torch
    import torch.distributed as dist
    import torch.distributed.nn
    import torch.nn.functional as F

    if gather_with_grad:
        A_all = torch.cat(torch.distributed.nn.all_gather(a_loc), dim=0)
        B_all = torch.cat(torch.distributed.nn.all_gather(b_loc), dim=0)
    else:
        la = [torch.zeros_like(a_loc) for _ in range(world_size)]
        lb = [torch.zeros_like(b_loc) for _ in range(world_size)]
        dist.all_gather(la, a_loc)
        dist.all_gather(lb, b_loc)
        if not local_loss:
            la[rank] = a_loc
            lb[rank] = b_loc
        A_all = torch.cat(la, dim=0)
        B_all = torch.cat(lb, dim=0)
    if local_loss:
        z_ab = (logit_scale * a_loc) @ B_all.T
        z_ba = (logit_scale * b_loc) @ A_all.T
    else:
        z_ab = (logit_scale * A_all) @ B_all.T
        z_ba = z_ab.T
    m = z_ab.shape[0]
    target = torch.arange(m, device=a_loc.device, dtype=torch.long)
    if local_loss:
        target = target + m * rank
    return (F.cross_entropy(z_ab, target) + F.cross_entropy(z_ba, target)) / 2


# ----------------------------------------------------------------------------------------------
# deterministic synthetic inputs: defined in tools/synthetic.py (bench.py's product arm must not
# import the oracle), re-exported here for the tests
# ----------------------------------------------------------------------------------------------
from tools.synthetic import synthetic_pair  # noqa: E402,F401


# ----------------------------------------------------------------------------------------------
# SigLipLoss (src/models/components/loss.py:204-311) - numpy float64 closed form
# ----------------------------------------------------------------------------------------------
@dataclass
class SigLipResult:
    loss: float          # value returned on this rank (loss.py:256-311)
    dA: np.ndarray       # gradient w.r.t. this rank's first positional feature tensor
    dB: np.ndarray       # gradient w.r.t. this rank's second positional feature tensor (ring backward: all ranks' losses)
    dscale: float = 0.0  # gradient of THIS rank's loss w.r.t. logit_scale (times its upstream gradient)
    dbias: float = 0.0   # ... w.r.t. logit_bias


def siglip_closed_form(A_all: np.ndarray, B_all: np.ndarray, scale: float, bias: float = 0.0, *, rank: int = 0,
                       world_size: int = 1, grad_outputs: Optional[np.ndarray] = None) -> SigLipResult:
    """Rank ``rank`` scores its own rows against the second operand of EVERY rank - its own block with
    labels 2I - 1 (loss.py:256), every other block with labels -1 (``negative_only``, loss.py:272-309;
    the uni- and bidirectional rings visit the same blocks) - and divides each block's sum by the local
    batch (loss.py:253).  The autograd of the neighbour exchanges (loss.py:169-201) returns to a rank
    the gradient of ITS second operand from every rank's loss, each times that rank's upstream
    gradient."""
    N, W = A_all.shape[0], world_size
    n = N // W
    g = np.ones(W) if grad_outputs is None else np.asarray(grad_outputs, dtype=np.float64)
    Z = scale * (A_all @ B_all.T) + bias
    Y = -np.ones((N, N))
    Y[np.arange(N), np.arange(N)] = 1.0
    M = -Y * Z
    lossmat = np.maximum(M, 0.0) + np.log1p(np.exp(-np.abs(M)))            # -logsigmoid(Y Z) = softplus(-Y Z)
    rows = slice(rank * n, (rank + 1) * n)
    dZ = (1.0 / (1.0 + np.exp(-Z)) - np.eye(N)) / n                         # d loss_r / d z_ij for i in rank r
    dZ = dZ * np.repeat(g, n)[:, None]
    return SigLipResult(loss=float(lossmat[rows].sum() / n), dA=scale * (dZ[rows] @ B_all), dB=scale * (dZ[:, rows].T @ A_all),
                        dscale=float((dZ[rows] * (A_all[rows] @ B_all.T)).sum()), dbias=float(dZ[rows].sum()))


# ----------------------------------------------------------------------------------------------
# all ranks at once (torch float64, any device): the same closed form as clip_loss_closed_form, arranged so that the
# N x N matrices are formed once for all W ranks - for the multi-GPU parity tests at sizes where W x 4 numpy passes
# over N x N doubles would take minutes (N = 16384: the 8-rank fused-gather shape).  Pinned against
# clip_loss_closed_form rank by rank in tests/test_oracle_golden.py.
# ----------------------------------------------------------------------------------------------
def clip_all_ranks_closed_form(A_all, B_all, scale: float, *, world_size: int, local_loss: bool, gather_with_grad: bool,
                               grad_outputs=None, device="cpu"):
    """-> dict(loss[W], dA[N, d], dB[N, d], dscale[W]) in float64 torch tensors (rows of rank r: [r n, (r + 1) n)).
    Conventions: see clip_loss_closed_form (loss.py:19-46, 85-114)."""
    import torch
    A = torch.as_tensor(A_all).to(device=device, dtype=torch.float64)
    B = torch.as_tensor(B_all).to(device=device, dtype=torch.float64)
    N, W = A.shape[0], int(world_size)
    n = N // W
    g = torch.ones(W, dtype=torch.float64, device=device) if grad_outputs is None else torch.as_tensor(grad_outputs).to(device=device, dtype=torch.float64)
    s = float(scale)
    dot = A @ B.T
    Z = s * dot
    rl = torch.logsumexp(Z, dim=1)
    cl = torch.logsumexp(Z, dim=0)
    diag = torch.diagonal(Z).clone()
    P = torch.exp(Z - rl[:, None])
    Q = torch.exp(Z - cl[None, :])
    idx = torch.arange(N, device=device)
    P[idx, idx] -= 1.0                        # P - I
    Q[idx, idx] -= 1.0                        # Q - I
    go = g.repeat_interleave(n)               # upstream gradient of the rank owning each row / column
    if W == 1 or not local_loss:
        Lg = 0.5 * ((rl - diag).mean() + (cl - diag).mean())
        loss = Lg.repeat(W)
        dZ = (P + Q) / (2 * N)                                   # dL_g / dZ
        unit_ds = (dZ * dot).sum()
        dA_u, dB_u = s * (dZ @ B), s * (dZ.T @ A)
        w = g.sum().repeat(N) if (W > 1 and gather_with_grad) else go
        dA, dB = w[:, None] * dA_u, w[:, None] * dB_u
        dscale = g * unit_ds
    else:
        per_row = 0.5 * ((rl - diag) + (cl - diag))
        loss = per_row.view(W, n).mean(dim=1)
        if gather_with_grad:
            gz = (go[:, None] * P + go[None, :] * Q) / (2 * n)
            dA, dB = s * (gz @ B), s * (gz.T @ A)
        else:                                                      # query side only
            dA = s * ((go[:, None] * P / (2 * n)) @ B)
            dB = s * ((go[None, :] * Q / (2 * n)).T @ A)
        # d value_r / d scale: rows of rank r through P, columns of rank r through Q
        rows_p = (P * dot).sum(dim=1).view(W, n).sum(dim=1)
        cols_q = (Q * dot).sum(dim=0).view(W, n).sum(dim=1)
        dscale = g * (rows_p + cols_q) / (2 * n)
    return dict(loss=loss, dA=dA, dB=dB, dscale=dscale)
