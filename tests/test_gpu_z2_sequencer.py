"""GPU: ClipLoss(host_sequencer=True) - each phase of the step enqueued from one C call
(oneprot_b200/csrc/clip_sequence.cu) - against the kernel-by-kernel Python host.  Both paths launch
the same kernels with the same arguments (tests/test_sequencer_cpu.py compares the launch traces), so
on one GPU the results must be BIT-identical; across GPUs the NVLS reductions are compared to 1e-6."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import clip_oracle as oc
from tests.helpers import cosine, rel_err

pytestmark = pytest.mark.gpu


def _step(a, b, scale=1.0, need=(True, True), **kw):
    from oneprot_b200 import ClipLoss, kernels
    A = a.cuda().requires_grad_(need[0])
    B = b.cuda().requires_grad_(need[1])
    m = ClipLoss(loss_dtype=torch.float32, **kw)
    kernels.launch_count_reset()
    loss = m(A, B, scale)
    loss.backward()
    torch.cuda.synchronize()
    m.check_last_call()
    return loss.detach(), A.grad, B.grad, kernels.launch_count()


@pytest.mark.parametrize("n,d,panel_rows", [(1024, 256, None), (300, 72 + 8, None), (25, 64, None), (2048, 512, 768),
                                            (4096, 1024, None)])
@pytest.mark.parametrize("need", [(True, True), (True, False), (False, True)])
def test_single_gpu_bit_identical_to_python_host(n, d, panel_rows, need):
    a, b = oc.synthetic_pair(n, d, seed=31 + n)
    kw = {}
    if panel_rows:
        kw["panel_bytes"] = 2 * ((n + 63) // 64 * 64) * panel_rows
        kw["keep_exp"] = False                       # several dL/dZ panels: the recompute backward
    l0, ga0, gb0, k0 = _step(a, b, need=need, host_sequencer=False, **kw)
    l1, ga1, gb1, k1 = _step(a, b, need=need, host_sequencer=True, **kw)
    assert l0.item() == l1.item()
    assert k0 == k1 and k1 > 0                       # same number of kernel launches
    for g0, g1 in ((ga0, ga1), (gb0, gb1)):
        assert (g0 is None) == (g1 is None)
        if g0 is not None:
            assert torch.equal(g0, g1)


def test_sequencer_matches_oracle_and_tensor_scale():
    """Also with a (non-differentiable) tensor logit_scale as in test_step (oneprot_module.py:142)."""
    n, d = 768, 128
    a, b = oc.synthetic_pair(n, d, seed=5, temperature_into_b=False)
    s = torch.tensor(1 / 0.07, device="cuda")
    ref = oc.clip_loss_closed_form(a.double().numpy(), b.double().numpy(), 1 / 0.07)
    l1, ga1, gb1, _ = _step(a, b, scale=s, host_sequencer=True)
    assert rel_err(l1.item(), ref.loss) < 1e-3
    assert cosine(ga1.float().cpu().numpy(), ref.dA) >= 0.9999 and cosine(gb1.float().cpu().numpy(), ref.dB) >= 0.9999


def test_sequencer_forward_without_grad():
    from oneprot_b200 import ClipLoss
    a, b = oc.synthetic_pair(512, 64, seed=9)
    with torch.no_grad():
        l0 = ClipLoss(loss_dtype=torch.float32)(a.cuda(), b.cuda())
        l1 = ClipLoss(loss_dtype=torch.float32, host_sequencer=True)(a.cuda(), b.cuda())
    assert l0.item() == l1.item()


# ---- multi-GPU (NVLS provider) ------------------------------------------------------------------
def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _worker(rank, world, port, n, d, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oneprot_b200 import ClipLoss
    from oneprot_b200.clip_loss import _get_comm
    a, b = oc.synthetic_pair(n, d, seed=77, rank=rank)
    rec = {"provider": None}
    for ll, gwg in ((False, True), (False, False), (True, True)):
        for seq in (False, True):
            for rep in range(2):          # twice: both parities of the double-buffered workspace
                A = a.cuda().requires_grad_(True)
                B = b.cuda().requires_grad_(True)
                m = ClipLoss(local_loss=ll, gather_with_grad=gwg, cache_labels=True, rank=rank, world_size=world,
                             loss_dtype=torch.float32, host_sequencer=seq,
                             panel_bytes=(5 << 28) if n != 1024 else 2 * (world * n) * 384, keep_exp=n != 1024)
                loss = m(A, B)
                (loss * (1.0 + 0.25 * rank)).backward()
                torch.cuda.synchronize()
                m.check_last_call()
            rec[(ll, gwg, seq)] = (loss.item(), A.grad.float().cpu(), B.grad.float().cpu())
    rec["provider"] = _get_comm(world, rank, None, torch.device("cuda", rank)).name
    results[rank] = rec
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("n,d", [(1024, 256), (512, 128)])
def test_multi_gpu_sequencer_equals_python_host(n, d):
    world = min(_ngpu(), 8)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, 29871 + n % 5, n, d, results), nprocs=world, join=True)
    for r in range(world):
        rec = results[r]
        if rec["provider"] != "nvls":
            pytest.skip("NVLS exchange provider unavailable: the sequencer stays on the Python path")
        for ll, gwg in ((False, True), (False, False), (True, True)):
            l0, ga0, gb0 = rec[(ll, gwg, False)]
            l1, ga1, gb1 = rec[(ll, gwg, True)]
            assert rel_err(l1, l0) < 1e-6, (r, ll, gwg)
            assert cosine(ga1.numpy(), ga0.numpy()) >= 0.999999 and cosine(gb1.numpy(), gb0.numpy()) >= 0.999999
            assert abs(ga1.norm().item() / ga0.norm().item() - 1) < 1e-5
