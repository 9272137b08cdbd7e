"""B200-native ``RetrievalMetric`` (reference: src/models/components/retrieval_metric.py:25-102).

Same ``update(preds, target)`` / ``compute()`` / ``reset()`` surface and the same result keys
(``{seq_to_mod,mod_to_seq}_median_rank`` and ``..._R@k``).  The reference materialises the N x N
similarity matrix on the CPU and argsorts it (retrieval_metric.py:89-97); here the rank of each label
is counted inside the epilogue of the logits tensor-core mainloop (``oneprot_retrieval_ranks``):
rank_i = #{ j != i : <s_i, m_j> > <s_i, m_i> }, which equals the argsort position whenever the row
has no exact ties with its label.  When torchmetrics is installed the class is a ``Metric`` with the
reference's list states (``dist_reduce_fx="cat"``), otherwise a plain accumulator.
"""
from __future__ import annotations

from typing import Any, List

import numpy as np
import torch

from . import kernels as _cuda_kernels
from .clip_loss import _prep_side

_KERNELS = _cuda_kernels

try:  # optional dependency of the reference
    from torchmetrics.metric import Metric as _Base
    from torchmetrics.utilities.data import dim_zero_cat
    _HAS_TM = True
except Exception:  # pragma: no cover
    _Base = object
    _HAS_TM = False

    def dim_zero_cat(x):
        return torch.cat(list(x), dim=0) if isinstance(x, (list, tuple)) else x


def retrieval_ranks(sequence_outputs: torch.Tensor, modality_outputs: torch.Tensor):
    """(rank_seq_to_mod, rank_mod_to_seq) as int64 CPU tensors of length N."""
    K = _KERNELS
    S, d, _ = _prep_side(sequence_outputs, 0)
    M, _, _ = _prep_side(modality_outputs, 1)
    N = S.shape[0]
    dev = S.device
    diag = torch.empty(N, dtype=torch.float32, device=dev)
    stats = torch.zeros(4, dtype=torch.float32, device=dev)
    K.rowstats(S, M, 0, diag, stats)
    r_s2m = torch.empty(N, dtype=torch.float32, device=dev)
    r_m2s = torch.empty(N, dtype=torch.float32, device=dev)
    K.retrieval_ranks(S, M, diag, r_s2m, r_m2s)
    return r_s2m.round().long().cpu(), r_m2s.round().long().cpu()


class RetrievalMetric(_Base):
    is_differentiable = False
    higher_is_better = True
    full_state_update = False

    def __init__(self, k: list = [1, 10, 100], **kwargs: Any) -> None:
        if _HAS_TM:
            super().__init__(**kwargs)
            self.add_state("preds", default=[], dist_reduce_fx="cat")
            self.add_state("target", default=[], dist_reduce_fx="cat")
        else:
            self.preds: List[torch.Tensor] = []
            self.target: List[torch.Tensor] = []
        self.k = k

    def update(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        self.preds.append(preds)
        self.target.append(target)

    def reset(self) -> None:
        if _HAS_TM:
            super().reset()
        else:
            self.preds, self.target = [], []

    def compute(self) -> dict:
        sequence_outputs = dim_zero_cat(self.preds)
        modality_outputs = dim_zero_cat(self.target)
        r_s2m, r_m2s = retrieval_ranks(sequence_outputs.detach(), modality_outputs.detach())
        metrics = {}
        for name, ranks in (("seq_to_mod", r_s2m.numpy()), ("mod_to_seq", r_m2s.numpy())):
            metrics[f"{name}_median_rank"] = np.floor(np.median(ranks)) + 1     # retrieval_metric.py:98
            for k in self.k:
                metrics[f"{name}_R@{k}"] = np.mean(ranks < k)                   # retrieval_metric.py:99-100
        return metrics
