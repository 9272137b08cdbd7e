"""Multi-GPU (NCCL) parity of the sharded ClipLoss: every (local_loss, gather_with_grad) convention
against the float64 closed-form oracle, plus the 2-rank reference golden.  Skipped with < 2 GPUs."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import clip_oracle as oc
from tests.helpers import bf16_from_bits, cosine, load_golden, rel_err

pytestmark = pytest.mark.gpu


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _worker(rank, world, port, n, d, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oneprot_b200 import ClipLoss
    rec = {}
    if n == 0:   # golden case
        g = load_golden("clip_dist_w2_n12_d32.npz")
        a, b = bf16_from_bits(g[f"r{rank}_A_bf16"]), bf16_from_bits(g[f"r{rank}_B_bf16"])
        s, gout = float(g["scale"]), float(g["grad_outputs"][rank])
    else:
        a, b = oc.synthetic_pair(n, d, seed=77, rank=rank, temperature_into_b=False)
        s, gout = 1.0 / 0.07, 1.0 + 0.25 * rank
    rec["a"], rec["b"] = a.float().numpy(), b.float().numpy()
    for ll in (False, True):
        for gwg in (False, True):
            keep_exp = n != 1024            # n = 1024 runs the recompute backward over several dL/dZ panels
            A = a.cuda().requires_grad_(True)
            B = b.cuda().requires_grad_(True)
            ls = torch.tensor(s, device="cuda", requires_grad=True)
            m = ClipLoss(local_loss=ll, gather_with_grad=gwg, cache_labels=True, rank=rank, world_size=world,
                         loss_dtype=torch.float32, panel_bytes=(1 << 30) if n != 1024 else 512 * 2048 * 2,
                         keep_exp=keep_exp)
            loss = m(A, B, ls)
            (loss * gout).backward()
            m.check_last_call()
            rec[f"ll{int(ll)}_gwg{int(gwg)}"] = dict(loss=loss.item(), dA=A.grad.float().cpu().numpy(),
                                                     dB=B.grad.float().cpu().numpy(), ds=ls.grad.item())
    results[rank] = rec
    dist.barrier()
    dist.destroy_process_group()


def _run(world, n, d, port):
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, d, results), nprocs=world, join=True)
    return results


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_reference_golden():
    res = _run(2, 0, 0, 29801)
    g = load_golden("clip_dist_w2_n12_d32.npz")
    for r in range(2):
        for tag in ("ll0_gwg0", "ll0_gwg1", "ll1_gwg0", "ll1_gwg1"):
            got = res[r][tag]
            assert rel_err(got["loss"], g[f"r{r}_loss_{tag}"]) < 1e-3, (r, tag)
            for k in ("dA", "dB"):
                want = g[f"r{r}_{k}_{tag}"]
                assert cosine(got[k], want) >= 0.9999, (r, tag, k)
                assert abs(np.linalg.norm(got[k]) / np.linalg.norm(want) - 1) < 1e-2, (r, tag, k)
            assert rel_err(got["ds"], g[f"r{r}_dscale_{tag}"]) < 2e-2, (r, tag)


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("n,d", [(1024, 256), (300, 128)])
def test_multi_gpu_closed_form(n, d):
    world = min(_ngpu(), 8 if n == 300 else 2)
    res = _run(world, n, d, 29811 + n % 7)
    A_all = np.concatenate([res[r]["a"] for r in range(world)]).astype(np.float64)
    B_all = np.concatenate([res[r]["b"] for r in range(world)]).astype(np.float64)
    gouts = np.array([1.0 + 0.25 * r for r in range(world)])
    for r in range(world):
        for ll in (False, True):
            for gwg in (False, True):
                ref = oc.clip_loss_closed_form(A_all, B_all, 1.0 / 0.07, rank=r, world_size=world, local_loss=ll,
                                               gather_with_grad=gwg, grad_outputs=gouts)
                got = res[r][f"ll{int(ll)}_gwg{int(gwg)}"]
                assert rel_err(got["loss"], ref.loss) < 1e-3
                assert cosine(got["dA"], ref.dA) >= 0.9999 and cosine(got["dB"], ref.dB) >= 0.9999
                assert abs(np.linalg.norm(got["dA"]) / np.linalg.norm(ref.dA) - 1) < 1e-2
                assert abs(got["ds"] - ref.dscale) < 2e-2 * abs(ref.dscale) + 1e-5
