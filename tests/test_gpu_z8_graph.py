"""GPU: ClipLoss(graph=True) - CUDA-graph replay of the forward / backward launch sequences (world_size 1) -
against the eager path, bit for bit.  Green on B200 since round 2."""
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


# ---- CUDA-graph replay over the static NVLS provider, world_size > 1 ------------------------------------------
def _graph_worker(rank, world, port, results):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oneprot_b200 import ClipLoss
    rec = {}
    n, d = 512, 128
    for ll, gwg in ((False, True), (True, True)):
        eager = ClipLoss(local_loss=ll, gather_with_grad=gwg, rank=rank, world_size=world, loss_dtype=torch.float32)
        graphed = ClipLoss(local_loss=ll, gather_with_grad=gwg, rank=rank, world_size=world, loss_dtype=torch.float32, graph=True)
        for step in range(4):                      # capture at step 0, three replays; new data every step
            a, b = oc.synthetic_pair(n, d, seed=100 + step, rank=rank, temperature_into_b=False)
            out = []
            for m in (eager, graphed):
                A = a.cuda().requires_grad_(True)
                B = b.cuda().requires_grad_(True)
                loss = m(A, B, 1.0 / 0.07)
                (loss * (1.0 + 0.25 * rank)).backward()
                torch.cuda.synchronize()
                out.append((loss.item(), A.grad.float().cpu().numpy(), B.grad.float().cpu().numpy()))
            rec[(ll, gwg, step)] = out
    results[rank] = rec
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif((torch.cuda.device_count() if torch.cuda.is_available() else 0) < 2, reason="needs >= 2 GPUs")
def test_multi_gpu_graph_replay_matches_eager():
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_graph_worker, args=(world, 29883, results), nprocs=world, join=True)
    for r in range(world):
        for key, (eager, graphed) in results[r].items():
            assert rel_err(graphed[0], eager[0]) < 1e-5, (r, key)
            for k in (1, 2):
                assert cosine(graphed[k], eager[k]) > 0.99999, (r, key, k)
                assert abs(np.linalg.norm(graphed[k]) / np.linalg.norm(eager[k]) - 1) < 2e-3, (r, key, k)
