"""PinnedPairPrefetcher (oneprot_b200/prefetch.py): pairs staged on the copy stream under the previous
step's kernels give bit-identical losses and gradients to pairs copied on the compute stream."""
import pytest
import torch

from oracle import clip_oracle as oc

pytestmark = pytest.mark.gpu


def _direct(a, b):
    from oneprot_b200 import ClipLoss
    A = a.cuda().requires_grad_(True)
    B = b.cuda().requires_grad_(True)
    loss = ClipLoss(loss_dtype=torch.float32)(A, B)
    loss.backward()
    return loss.detach().clone(), A.grad.clone(), B.grad.clone()


@pytest.mark.parametrize("slots", [2, 3])
def test_prefetched_steps_equal_direct_steps(slots):
    from oneprot_b200 import ClipLoss
    from oneprot_b200.prefetch import PinnedPairPrefetcher
    n, d, steps = 1024, 256, 6
    pairs = [tuple(t.pin_memory() for t in oc.synthetic_pair(n, d, seed=100 + k)) for k in range(steps)]
    want = [_direct(a, b) for a, b in pairs]
    pf = PinnedPairPrefetcher("cuda", slots=slots)
    clip = ClipLoss(loss_dtype=torch.float32)
    pf.submit(*pairs[0])
    got = []
    for k in range(steps):
        A, B = pf.next()
        if k + 1 < steps:
            pf.submit(*pairs[k + 1])          # flies under this step's kernels
        A.requires_grad_(True); B.requires_grad_(True)
        loss = clip(A, B)
        loss.backward()
        got.append((loss.detach().clone(), A.grad.clone(), B.grad.clone()))
    torch.cuda.synchronize()
    for (l0, ga0, gb0), (l1, ga1, gb1) in zip(want, got):
        assert l0.item() == l1.item()
        assert torch.equal(ga0, ga1) and torch.equal(gb0, gb1)


def test_prefetcher_ring_discipline():
    from oneprot_b200.prefetch import PinnedPairPrefetcher
    a, b = (t.pin_memory() for t in oc.synthetic_pair(64, 32, seed=1))
    pf = PinnedPairPrefetcher("cuda")
    with pytest.raises(RuntimeError):
        pf.next()                               # nothing submitted
    with pytest.raises(ValueError):
        pf.submit(a.clone(), b)                 # pageable host memory
    pf.submit(a, b); pf.submit(a, b)
    with pytest.raises(RuntimeError):
        pf.submit(a, b)                         # both slots waiting to be consumed
    A0, _ = pf.next()
    with pytest.raises(RuntimeError):
        pf.submit(a, b)                         # slot 0 is still owned by the consumer
    A1, _ = pf.next()                           # releases slot 0
    pf.submit(a, b)
    A2, _ = pf.next()
    torch.cuda.synchronize()
    assert A2.data_ptr() == A0.data_ptr() != A1.data_ptr()
    assert torch.equal(A2.cpu(), a)
