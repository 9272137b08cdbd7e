"""GPU: SigLipLoss (oneprot_b200/siglip_loss.py; clip_s_kernel<SFWD> / <SDZ>) against the
reference-generated golden fixtures and the float64 closed form.  Green on B200 since round 2.
Tolerances as for ClipLoss: loss <= 1e-3 relative for bf16 features (against fp64 on the same
bf16-valued inputs), <= 1e-5 for fp32 features, gradient cosine >= 0.9999 (+ norm within 1 %)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import clip_oracle as oc
from tests.helpers import bf16_from_bits, cosine, load_golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["train", "paper", "odd"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("keep", [False, True])          # True: kept sigma panel (clip_s_kernel<SFWD_K>), no panel kernel in the backward
def test_siglip_golden_single_gpu(tag, dtype, keep):
    from oneprot_b200.siglip_loss import SigLipLoss
    g = load_golden("siglip_single.npz")
    a, b = bf16_from_bits(g[f"{tag}_A_bf16"]), bf16_from_bits(g[f"{tag}_B_bf16"])
    A = a.to(dtype).cuda().requires_grad_(True)
    B = b.to(dtype).cuda().requires_grad_(True)
    bias = float(g[f"{tag}_bias"]) if bool(g[f"{tag}_has_bias"]) else None
    m = SigLipLoss(keep_exp=keep)
    st = torch.tensor(float(g[f"{tag}_scale"]), device="cuda", requires_grad=True)
    bt = None if bias is None else torch.tensor(bias, device="cuda", requires_grad=True)
    loss = m(A, B, st, bt)
    assert loss.dtype == dtype
    loss.backward()
    assert abs(st.grad.item() - float(g[f"{tag}_dscale"])) < 2e-2 * abs(float(g[f"{tag}_dscale"])) + 1e-6
    if bt is not None:
        assert abs(bt.grad.item() - float(g[f"{tag}_dbias"])) < 2e-3 * abs(float(g[f"{tag}_dbias"])) + 1e-6
    assert rel_err(m.last_loss_fp32.item(), g[f"{tag}_loss"]) < (1e-3 if dtype == torch.bfloat16 else 1e-5)
    for got, want in ((A.grad, g[f"{tag}_dA"]), (B.grad, g[f"{tag}_dB"])):
        assert cosine(got.double().cpu().numpy(), want) >= 0.9999
        assert abs(np.linalg.norm(got.double().cpu().numpy()) / np.linalg.norm(want) - 1) < 1e-2


@pytest.mark.parametrize("keep", [False, True])
@pytest.mark.parametrize("n,d,scale,bias,panel_rows", [(1000, 256, 10.0, -10.0, None), (2048, 1024, 1.0, None, 768),
                                                        (300, 64, 25.0, 3.0, None)])
def test_siglip_vs_closed_form(n, d, scale, bias, panel_rows, keep):
    from oneprot_b200.siglip_loss import SigLipLoss
    a, b = oc.synthetic_pair(n, d, seed=n + d, temperature_into_b=(bias is None))
    ref = oc.siglip_closed_form(a.double().numpy(), b.double().numpy(), scale, 0.0 if bias is None else bias)
    kw = {} if panel_rows is None else dict(panel_bytes=2 * ((n + 63) // 64 * 64) * panel_rows)
    outs = []
    for _ in range(2):                                 # twice: bit-reproducible
        A = a.cuda().requires_grad_(True)
        B = b.cuda().requires_grad_(True)
        m = SigLipLoss(loss_dtype=torch.float32, keep_exp=keep, **kw)
        loss = m(A, B, scale, None if bias is None else torch.tensor(bias))
        loss.backward()
        outs.append((loss.item(), A.grad.clone(), B.grad.clone()))
    assert outs[0][0] == outs[1][0] and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    assert rel_err(outs[0][0], ref.loss) < 1e-3
    assert cosine(outs[0][1].float().cpu().numpy(), ref.dA) >= 0.9999 and cosine(outs[0][2].float().cpu().numpy(), ref.dB) >= 0.9999
    assert abs(np.linalg.norm(outs[0][2].float().cpu().numpy()) / np.linalg.norm(ref.dB) - 1) < 1e-2


def test_siglip_extreme_logits_do_not_overflow():
    """|z| of several hundred: softplus / sigmoid must stay finite (2^x would overflow fp32 at x > 128)."""
    from oneprot_b200.siglip_loss import SigLipLoss
    a, b = oc.synthetic_pair(256, 64, seed=3, temperature_into_b=False)
    ref = oc.siglip_closed_form(a.double().numpy(), b.double().numpy(), 400.0, -50.0)
    A = a.cuda().requires_grad_(True)
    B = b.cuda().requires_grad_(True)
    loss = SigLipLoss(loss_dtype=torch.float32)(A, B, 400.0, -50.0)
    loss.backward()
    assert np.isfinite(loss.item()) and torch.isfinite(A.grad).all() and torch.isfinite(B.grad).all()
    assert rel_err(loss.item(), ref.loss) < 5e-3
    assert cosine(A.grad.float().cpu().numpy(), ref.dA) >= 0.999


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _worker(rank, world, port, n, d, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oneprot_b200.siglip_loss import SigLipLoss
    a, b = oc.synthetic_pair(n, d, seed=55, rank=rank, temperature_into_b=False)
    A = a.cuda().requires_grad_(True)
    B = b.cuda().requires_grad_(True)
    loss = SigLipLoss(rank=rank, world_size=world, loss_dtype=torch.float32)(A, B, 10.0, -10.0)
    (loss * (1.0 + 0.25 * rank)).backward()
    torch.cuda.synchronize()
    results[rank] = (a.float().numpy(), b.float().numpy(), loss.item(), A.grad.float().cpu().numpy(), B.grad.float().cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_siglip_multi_gpu_closed_form():
    world, n, d = min(_ngpu(), 4), 512, 128
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, 29881, n, d, results), nprocs=world, join=True)
    A_all = np.concatenate([results[r][0] for r in range(world)]).astype(np.float64)
    B_all = np.concatenate([results[r][1] for r in range(world)]).astype(np.float64)
    gouts = np.array([1.0 + 0.25 * r for r in range(world)])
    for r in range(world):
        ref = oc.siglip_closed_form(A_all, B_all, 10.0, -10.0, rank=r, world_size=world, grad_outputs=gouts)
        _, _, loss, dA, dB = results[r]
        assert rel_err(loss, ref.loss) < 1e-3
        assert cosine(dA, ref.dA) >= 0.9999 and cosine(dB, ref.dB) >= 0.9999
