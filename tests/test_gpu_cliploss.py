"""GPU parity of the drop-in ClipLoss (CUDA path through the C ABI) against
  * the reference-generated golden fixtures (tests/golden),
  * the float64 closed-form oracle on seeded synthetic embeddings,
  * size-independent properties at BASELINE.json's full sizes.
Tolerances are the ones BASELINE.json's north_star states: loss <= 1e-3 relative for bf16 inputs
(judged against fp64 on the same bf16-valued inputs), <= 1e-5 for fp32 inputs,
gradient cosine >= 0.9999."""
import glob
import math
import os

import numpy as np
import pytest
import torch

from oracle import clip_oracle as oc
from tests.helpers import GOLDEN, bf16_from_bits, cosine, load_golden, rel_err

pytestmark = pytest.mark.gpu

BF16_LOSS_RTOL = 1e-3
FP32_LOSS_RTOL = 1e-5
GRAD_COS = 0.9999

SINGLE = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "clip_single_*.npz")))


def _loss_mod(**kw):
    from oneprot_b200 import ClipLoss
    return ClipLoss(**kw)


@pytest.mark.parametrize("keep_exp", [True, False])     # stored-exponentials backward (default) / recompute backward
@pytest.mark.parametrize("name", SINGLE)
def test_golden_single_rank(name, keep_exp):
    g = load_golden(name)
    A = bf16_from_bits(g["A_bf16"]).cuda().requires_grad_(True)
    B = bf16_from_bits(g["B_bf16"]).cuda().requires_grad_(True)
    is_t = bool(g["scale_is_tensor"])
    ls = torch.tensor(float(g["scale"]), device="cuda", requires_grad=True) if is_t else float(g["scale"])
    m = _loss_mod(loss_dtype=torch.float32, keep_exp=keep_exp)
    loss = m(A, B, ls)
    loss.backward()
    torch.cuda.synchronize()
    assert rel_err(loss.item(), g["loss_f64"]) < BF16_LOSS_RTOL
    keep = g["dA_f64"].shape[0]
    assert cosine(A.grad[:keep].float().cpu().numpy(), g["dA_f64"]) >= GRAD_COS
    assert cosine(B.grad[:keep].float().cpu().numpy(), g["dB_f64"]) >= GRAD_COS
    na, nw = np.linalg.norm(A.grad[:keep].float().cpu().numpy()), np.linalg.norm(g["dA_f64"])
    assert abs(na / nw - 1) < 1e-2
    if is_t:
        assert rel_err(ls.grad.item(), g["dscale_f64"]) < 2e-2
    m.check_last_call()
    # reference return dtype (bf16 in -> bf16 out); value within one bf16 ulp of the fp64 truth
    lb = _loss_mod()(A.detach(), B.detach(), float(g["scale"]))
    assert lb.dtype == torch.bfloat16 and lb.dim() == 0
    assert abs(lb.item() - float(g["loss_f64"])) <= 2 ** (math.floor(math.log2(abs(float(g["loss_f64"])))) - 7)


@pytest.mark.parametrize("n,d,corr,t_in_b", [(64, 1024, True, True), (1000, 1024, True, True), (2048, 1024, False, True),
                                             (96, 512, True, False), (3000, 256, True, True)])
@pytest.mark.parametrize("keep_exp", [True, False])
def test_synthetic_vs_closed_form(n, d, corr, t_in_b, keep_exp):
    a, b = oc.synthetic_pair(n, d, seed=1234, pair_id=1, rank=0, correlated=corr, temperature_into_b=t_in_b)
    s = 1.0 if t_in_b else 1.0 / 0.07
    ref = oc.clip_loss_closed_form(a.double().numpy(), b.double().numpy(), s)
    A = a.cuda().requires_grad_(True)
    B = b.cuda().requires_grad_(True)
    ls = torch.tensor(s, device="cuda", requires_grad=True)
    m = _loss_mod(loss_dtype=torch.float32, keep_exp=keep_exp)
    loss = m(A, B, ls)
    loss.backward()
    assert rel_err(loss.item(), ref.loss) < BF16_LOSS_RTOL
    assert cosine(A.grad.float().cpu().numpy(), ref.dA) >= GRAD_COS
    assert cosine(B.grad.float().cpu().numpy(), ref.dB) >= GRAD_COS
    assert rel_err(ls.grad.item(), ref.dscale) < 2e-2 or abs(ls.grad.item() - ref.dscale) < 1e-4
    m.check_last_call()


def test_multi_panel_backward_equals_single_panel():
    a, b = oc.synthetic_pair(1500, 512, seed=7)
    outs = []
    for pb in (1 << 30, 256 * 1536 * 2):
        A = a.cuda().requires_grad_(True)
        B = b.cuda().requires_grad_(True)
        _loss_mod(loss_dtype=torch.float32, panel_bytes=pb, keep_exp=False)(A, B).backward()
        outs.append((A.grad.float().cpu().numpy(), B.grad.float().cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0])          # dA rows are panel-independent: bit-exact
    assert cosine(outs[0][1], outs[1][1]) > 0.99999         # dB: fp32 accumulation across panels


def test_fp32_inputs_meet_fp32_tolerance():
    a, b = oc.synthetic_pair(512, 256, seed=11, dtype="fp32")
    ref = oc.clip_loss_closed_form(a.double().numpy(), b.double().numpy(), 1.0)
    A = a.cuda().requires_grad_(True)
    B = b.cuda().requires_grad_(True)
    loss = _loss_mod()(A, B)
    assert loss.dtype == torch.float32
    loss.backward()
    assert rel_err(loss.item(), ref.loss) < FP32_LOSS_RTOL
    assert A.grad.dtype == torch.float32
    assert cosine(A.grad.cpu().numpy(), ref.dA) >= GRAD_COS and cosine(B.grad.cpu().numpy(), ref.dB) >= GRAD_COS


def test_no_grad_and_noncontiguous_inputs():
    a, b = oc.synthetic_pair(300, 128, seed=3)
    A = torch.empty(300, 256, dtype=torch.bfloat16, device="cuda")[:, ::2]
    A.copy_(a)
    with torch.no_grad():
        l1 = _loss_mod(loss_dtype=torch.float32)(A, b.cuda())
    l2 = _loss_mod(loss_dtype=torch.float32)(a.cuda(), b.cuda())
    assert l1.item() == l2.item() and not l1.requires_grad


def test_properties_at_full_size():
    """N = 8192 x 1024 (BASELINE configs[1] shape): properties that need no N x N oracle."""
    n, d = 8192, 1024
    a, b = oc.synthetic_pair(n, d, seed=1234, correlated=False)
    A = a.cuda().requires_grad_(True)
    B = b.cuda().requires_grad_(True)
    m = _loss_mod(loss_dtype=torch.float32)
    loss = m(A, B)
    loss.backward()
    m.check_last_call()
    # (1) symmetry: swapping the operands leaves the value unchanged and swaps the gradients
    A2 = a.cuda().requires_grad_(True)
    B2 = b.cuda().requires_grad_(True)
    loss_sw = _loss_mod(loss_dtype=torch.float32)(B2, A2)
    loss_sw.backward()
    assert rel_err(loss_sw.item(), loss.item()) < 1e-5
    assert cosine(A2.grad.float().cpu().numpy(), A.grad.float().cpu().numpy()) > 0.9999
    # (2) softmax gradients sum to zero against any constant direction: sum_i dA_i is
    #     s * sum_ij dZ_ij b_j with sum_j dZ_ij = 0 only in expectation; instead use the exact
    #     identity <A, dA> = <B, dB> (both equal scale * dL/dscale)
    ta = (A.detach().float() * A.grad.float()).sum().item()
    tb = (B.detach().float() * B.grad.float()).sum().item()
    assert abs(ta - tb) < 2e-2 * max(abs(ta), abs(tb)) + 1e-4
    # (3) uncorrelated unit-norm anchors: loss is within a small margin of ln N
    assert abs(loss.item() - math.log(n)) < 0.35
    # (4) determinism: bit-identical on repeat
    A3 = a.cuda().requires_grad_(True)
    B3 = b.cuda().requires_grad_(True)
    l3 = _loss_mod(loss_dtype=torch.float32)(A3, B3)
    l3.backward()
    assert l3.item() == loss.item() and torch.equal(A3.grad, A.grad) and torch.equal(B3.grad, B.grad)


def test_full_size_vs_fp32_port_on_gpu():
    """N = 8192: compare with the torch port of the reference run in fp32 on the GPU (TF32 off)."""
    n, d = 8192, 1024
    a, b = oc.synthetic_pair(n, d, seed=99)
    A = a.cuda().requires_grad_(True)
    B = b.cuda().requires_grad_(True)
    _l = _loss_mod(loss_dtype=torch.float32)(A, B)
    _l.backward()
    torch.backends.cuda.matmul.allow_tf32 = False
    lp, dA, dB = oc.clip_loss_port_fwd_bwd(a.cuda().float(), b.cuda().float(), 1.0)
    assert rel_err(_l.item(), lp.item()) < BF16_LOSS_RTOL
    assert cosine(A.grad.float().cpu().numpy(), dA.cpu().numpy()) >= GRAD_COS
    assert cosine(B.grad.float().cpu().numpy(), dB.cpu().numpy()) >= GRAD_COS


@pytest.mark.parametrize("n,d,amp", [(256, 512, 1.0), (300, 64, 1.5), (1024, 1024, 1.0)])
def test_unnormalised_inputs_use_exact_max_reference(n, d, amp):
    """|c| max|a| max|b| >> 100: the Cauchy-Schwarz reference alone would underflow every term; the
    max pass supplies the exact maximum logit and the result still matches the float64 oracle."""
    g = torch.Generator().manual_seed(0)
    a = (amp * torch.randn(n, d, generator=g)).to(torch.bfloat16)
    b = (amp * torch.randn(n, d, generator=g)).to(torch.bfloat16)
    ref = oc.clip_loss_closed_form(a.double().numpy(), b.double().numpy(), 1.0)
    A = a.cuda().requires_grad_(True)
    B = b.cuda().requires_grad_(True)
    m = _loss_mod(loss_dtype=torch.float32)
    loss = m(A, B, 1.0)
    loss.backward()
    m.check_last_call()
    assert rel_err(loss.item(), ref.loss) < BF16_LOSS_RTOL
    assert cosine(A.grad.float().cpu().numpy(), ref.dA) >= GRAD_COS
    assert cosine(B.grad.float().cpu().numpy(), ref.dB) >= GRAD_COS


def test_hazard_flag_when_row_maxima_are_hundreds_of_bits_apart():
    """Beyond the validated window (one row's best logit is > 2^200 below the global maximum) the
    device flag is raised instead of returning a silently wrong value."""
    g = torch.Generator().manual_seed(0)
    a = torch.randn(256, 64, generator=g)
    b = torch.randn(256, 64, generator=g)
    a[:128] *= 60.0          # logits of the first rows are ~60x larger than those of the last rows
    b[:128] *= 60.0
    m = _loss_mod(loss_dtype=torch.float32)
    m(a.to(torch.bfloat16).cuda(), b.to(torch.bfloat16).cuda(), 1.0)
    with pytest.raises(FloatingPointError):
        m.check_last_call()


def test_odd_feature_dim_and_fp16_inputs():
    g = torch.Generator().manual_seed(4)
    a = torch.nn.functional.normalize(torch.randn(70, 13, generator=g), dim=-1)
    b = torch.nn.functional.normalize(torch.randn(70, 13, generator=g), dim=-1) * 8
    for dt, tol in ((torch.bfloat16, BF16_LOSS_RTOL), (torch.float16, 1e-4)):
        A = a.to(dt).cuda().requires_grad_(True)
        B = b.to(dt).cuda().requires_grad_(True)
        ref = oc.clip_loss_closed_form(A.detach().double().cpu().numpy(), B.detach().double().cpu().numpy(), 1.0)
        loss = _loss_mod(loss_dtype=torch.float32)(A, B)
        loss.backward()
        assert rel_err(loss.item(), ref.loss) < tol
        assert A.grad.shape == (70, 13) and A.grad.dtype == dt
        assert cosine(A.grad.float().cpu().numpy(), ref.dA) >= GRAD_COS


def test_epilogue_modules_match_golden():
    from oneprot_b200 import LearnableLogitScaling, Normalize
    g = load_golden("epilogue_normalize_scale.npz")
    x = bf16_from_bits(g["x_bf16"]).cuda().float().requires_grad_(True)
    gy = bf16_from_bits(g["gy_bf16"]).cuda().float()
    y = Normalize(dim=-1)(x)
    y.backward(gy)
    assert np.allclose(y.detach().cpu().numpy(), g["y_f64"], rtol=1e-5, atol=1e-6)
    assert np.allclose(x.grad.cpu().numpy(), g["gx_f64"], rtol=1e-4, atol=1e-5)
    sc = LearnableLogitScaling(logit_scale_init=1 / 0.07, learnable=True).cuda()
    ys = sc(y.detach())
    assert np.allclose(ys.detach().cpu().numpy(), g["ys_f32"], rtol=1e-5)
    big = LearnableLogitScaling(logit_scale_init=250.0, learnable=False).cuda()
    assert np.allclose(big(y.detach()).cpu().numpy(), g["yb_f32"], rtol=1e-5)
