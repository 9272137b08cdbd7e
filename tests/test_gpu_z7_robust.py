"""GPU: two-reference (robust) ClipLoss path - per-row / per-column references for inputs outside the
single-reference fp32 window - against the float64 oracle.  Green on B200 since round 2."""
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


@pytest.mark.parametrize("mode", ["always", "auto"])
def test_two_reference_path_on_wild_inputs(mode):
    """Row maxima thousands of bits apart (outside the single-reference window): robust modes must
    still match the float64 oracle."""
    g = torch.Generator().manual_seed(3)
    a = torch.randn(300, 128, generator=g)
    b = torch.randn(300, 128, generator=g)
    a[:150] *= 40.0
    b[:100] *= 25.0
    a, b = a.to(torch.bfloat16), b.to(torch.bfloat16)
    ref = oc.clip_loss_closed_form(a.double().numpy(), b.double().numpy(), 1.0)
    A = a.cuda().requires_grad_(True)
    B = b.cuda().requires_grad_(True)
    m = _loss_mod(loss_dtype=torch.float32, robust=mode)
    loss = m(A, B, 1.0)
    loss.backward()
    m.check_last_call()
    assert rel_err(loss.item(), ref.loss) < BF16_LOSS_RTOL
    assert cosine(A.grad.float().cpu().numpy(), ref.dA) >= GRAD_COS
    assert cosine(B.grad.float().cpu().numpy(), ref.dB) >= GRAD_COS


def test_two_reference_path_matches_default_on_normalised_inputs():
    a, b = oc.synthetic_pair(1000, 512, seed=21)
    outs = []
    for mode in ("off", "always"):
        A = a.cuda().requires_grad_(True)
        B = b.cuda().requires_grad_(True)
        loss = _loss_mod(loss_dtype=torch.float32, robust=mode)(A, B)
        loss.backward()
        outs.append((loss.item(), A.grad.float().cpu().numpy(), B.grad.float().cpu().numpy()))
    assert rel_err(outs[1][0], outs[0][0]) < 1e-5
    assert cosine(outs[0][1], outs[1][1]) > 0.9999 and cosine(outs[0][2], outs[1][2]) > 0.9999
