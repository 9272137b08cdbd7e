"""GPU: ClipLoss(graph=True) - CUDA-graph replay of the forward / backward launch sequences (world_size 1) -
against the eager path, bit for bit.  Not yet run on hardware."""
import math

import numpy as np
import pytest
import torch

from oracle import clip_oracle as oc
from tests.helpers import cosine, rel_err

pytestmark = pytest.mark.gpu

BF16_LOSS_RTOL = 1e-3
GRAD_COS = 0.9999


def _loss_mod(**kw):
    from oneprot_b200 import ClipLoss
    return ClipLoss(**kw)


def test_graph_mode_replays_match_eager():
    """ClipLoss(graph=True): captured forward / backward graphs give the eager results on new data."""
    outs = {}
    for graph in (False, True):
        m = _loss_mod(loss_dtype=torch.float32, graph=graph)
        res = []
        for seed in (31, 32, 33):
            a, b = oc.synthetic_pair(1000, 256, seed=seed)
            A = a.cuda().requires_grad_(True)
            B = b.cuda().requires_grad_(True)
            loss = m(A, B)
            loss.backward()
            res.append((loss.item(), A.grad.float().cpu().numpy(), B.grad.float().cpu().numpy()))
        outs[graph] = res
    for (l0, a0, b0), (l1, a1, b1) in zip(outs[False], outs[True]):
        assert l0 == l1
        assert np.array_equal(a0, a1) and np.array_equal(b0, b1)
