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


def test_stale_backward_is_refused():
    """Two forwards on one graphed module before the first backward: the static buffers hold the second call's state,
    so the first backward raises instead of returning the second call's gradients."""
    from oneprot_b200 import ClipLoss
    a, b = oc.synthetic_pair(256, 64, seed=3)
    m = ClipLoss(loss_dtype=torch.float32, graph=True)
    A1, B1 = a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    A2, B2 = torch.roll(a, 1, 0).cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    l1 = m(A1, B1)
    l2 = m(A2, B2)
    l2.backward()                      # the latest forward: fine
    assert A2.grad is not None
    with pytest.raises(RuntimeError, match="no longer the latest"):
        l1.backward()
