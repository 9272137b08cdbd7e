"""CPU: pins oracle/clip_oracle.py against fixtures generated from the unmodified reference
(oracle/make_golden.py).  The reference's own tests hold no golden vector for this path
(SURVEY.md section 4), so these reference-generated fixtures are the pin."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import clip_oracle as oc
from tests.helpers import GOLDEN, bf16_from_bits, cosine, load_golden, rel_err

SINGLE = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "clip_single_*.npz")))


@pytest.mark.parametrize("name", SINGLE)
def test_closed_form_matches_reference_single(name):
    g = load_golden(name)
    A = bf16_from_bits(g["A_bf16"]).double().numpy()
    B = bf16_from_bits(g["B_bf16"]).double().numpy()
    res = oc.clip_loss_closed_form(A, B, float(g["scale"]))
    assert rel_err(res.loss, g["loss_f64"]) < 1e-12
    keep = g["dA_f64"].shape[0]
    np.testing.assert_allclose(res.dA[:keep], g["dA_f64"], rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(res.dB[:keep], g["dB_f64"], rtol=2e-6, atol=1e-9)
    if bool(g["scale_is_tensor"]):
        assert rel_err(res.dscale, g["dscale_f64"]) < 1e-9


@pytest.mark.parametrize("name", SINGLE)
def test_port_matches_reference_single(name):
    g = load_golden(name)
    a = bf16_from_bits(g["A_bf16"])
    b = bf16_from_bits(g["B_bf16"])
    s = float(g["scale"])
    for tag, td, tol in (("f64", torch.float64, 1e-12), ("f32", torch.float32, 1e-6)):
        ls = torch.tensor(s, dtype=td) if bool(g["scale_is_tensor"]) else 1.0
        loss, dA, dB = oc.clip_loss_port_fwd_bwd(a.to(td), b.to(td), ls)
        assert rel_err(loss.item(), g[f"loss_{tag}"]) < tol
        if tag == "f64":
            keep = g["dA_f64"].shape[0]
            assert cosine(dA[:keep].numpy(), g["dA_f64"]) > 1 - 1e-10
            assert cosine(dB[:keep].numpy(), g["dB_f64"]) > 1 - 1e-10
    # the reference returns a bf16 scalar for bf16 inputs; the port does too
    ls = torch.tensor(s, dtype=torch.bfloat16) if bool(g["scale_is_tensor"]) else 1.0
    lb = oc.clip_loss_port(a, b, ls)
    assert lb.dtype == torch.bfloat16
    assert float(lb) == float(g["loss_bf16"])


def test_closed_form_matches_reference_distributed_conventions():
    g = load_golden("clip_dist_w2_n12_d32.npz")
    W = int(g["world"])
    A_all = np.concatenate([bf16_from_bits(g[f"r{r}_A_bf16"]).double().numpy() for r in range(W)])
    B_all = np.concatenate([bf16_from_bits(g[f"r{r}_B_bf16"]).double().numpy() for r in range(W)])
    for ll in (False, True):
        for gwg in (False, True):
            tag = f"ll{int(ll)}_gwg{int(gwg)}"
            for r in range(W):
                res = oc.clip_loss_closed_form(A_all, B_all, float(g["scale"]), rank=r, world_size=W,
                                               local_loss=ll, gather_with_grad=gwg,
                                               grad_outputs=g["grad_outputs"])
                assert rel_err(res.loss, g[f"r{r}_loss_{tag}"]) < 1e-12, tag
                np.testing.assert_allclose(res.dA, g[f"r{r}_dA_{tag}"], rtol=1e-9, atol=1e-13, err_msg=tag)
                np.testing.assert_allclose(res.dB, g[f"r{r}_dB_{tag}"], rtol=1e-9, atol=1e-13, err_msg=tag)
                assert rel_err(res.dscale, g[f"r{r}_dscale_{tag}"]) < 1e-9, tag


def test_epilogue_closed_forms():
    g = load_golden("epilogue_normalize_scale.npz")
    x = bf16_from_bits(g["x_bf16"]).double().numpy()
    gy = bf16_from_bits(g["gy_bf16"]).double().numpy()
    np.testing.assert_allclose(oc.normalize_closed_form(x), g["y_f64"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(oc.normalize_backward_closed_form(x, gy), g["gx_f64"], rtol=1e-9, atol=1e-12)
    y = g["y_f64"].astype(np.float32).astype(np.float64)
    ys, s = oc.logit_scaling_closed_form(y, float(g["log_logit_scale"]))
    np.testing.assert_allclose(ys, g["ys_f32"], rtol=1e-6)
    yb, sb = oc.logit_scaling_closed_form(y, float(g["log_logit_scale_big"]))
    assert sb == 100.0
    np.testing.assert_allclose(yb, g["yb_f32"], rtol=1e-6)


def test_synthetic_generator_is_deterministic():
    a1, b1 = oc.synthetic_pair(16, 32, seed=5, rank=1)
    a2, b2 = oc.synthetic_pair(16, 32, seed=5, rank=1)
    assert torch.equal(a1, a2) and torch.equal(b1, b2)
    assert a1.dtype == torch.bfloat16
    # unit-norm anchor, temperature folded into B
    assert abs(a1.float().norm(dim=-1).mean().item() - 1.0) < 1e-2
    assert abs(b1.float().norm(dim=-1).mean().item() - 1 / 0.07) < 0.2


def test_panel_port_is_the_row_restriction_of_the_port():
    a, b = oc.synthetic_pair(64, 32, seed=9, dtype="fp32")
    full_ab = torch.nn.functional.cross_entropy(a @ b.T, torch.arange(64), reduction="none")
    full_ba = torch.nn.functional.cross_entropy(b @ a.T, torch.arange(64), reduction="none")
    want = (full_ab[:16].mean() + full_ba[:16].mean()) / 2
    got = oc.clip_loss_port_panel(a, b, 16, 1.0)
    assert abs(got.item() - want.item()) < 1e-6
    assert abs(oc.clip_loss_port_panel(a, b, 64, 1.0).item() - oc.clip_loss_port(a, b, 1.0).item()) < 1e-6


def test_retrieval_metric_oracle_matches_reference_golden():
    """retrieval_metric_closed_form against the reference's RetrievalMetric (retrieval_metric.py:71-102;
    torchmetrics' Metric base replaced by a list-state stand-in when the fixture was generated)."""
    g = load_golden("retrieval_metric.npz")
    for tag in ("easy", "hard"):
        S = bf16_from_bits(g[f"{tag}_S_bf16"]).double().numpy()
        M = bf16_from_bits(g[f"{tag}_M_bf16"]).double().numpy()
        got = oc.retrieval_metric_closed_form(S, M)
        want = {k.split(":", 1)[1]: float(v) for k, v in g.items() if k.startswith(tag + ":")}
        assert set(got) == set(want)
        for k in want:
            assert float(got[k]) == want[k], (tag, k)


# ---- oracle/_ref: the unmodified reference classes (vendored by oracle/make_ref.py), when present ---------------
def _ref_or_skip():
    from oracle.make_ref import import_reference
    ref = import_reference()
    if ref is None:
        pytest.skip("oracle/_ref not built (python oracle/make_ref.py needs /root/reference)")
    return ref


@pytest.mark.parametrize("n,d,scale", [(256, 512, 1.0 / 0.07), (100, 72, 1.0), (33, 40, 5.0)])
def test_port_and_closed_form_match_the_vendored_reference(n, d, scale):
    """The port bench.py times and the closed form the GPU tests use, against the reference's own ClipLoss run here."""
    import torch
    loss_mod, _ = _ref_or_skip()
    g = torch.Generator().manual_seed(n)
    a = torch.nn.functional.normalize(torch.randn(n, d, generator=g, dtype=torch.float64), dim=-1)
    b = torch.nn.functional.normalize(a + 0.5 * torch.randn(n, d, generator=g, dtype=torch.float64), dim=-1)
    A, B = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    s = torch.tensor(scale, dtype=torch.float64, requires_grad=True)
    want = loss_mod.ClipLoss()(A, B, s)
    want.backward()
    A2, B2 = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    got = oc.clip_loss_port(A2, B2, scale)
    got.backward()
    assert abs(got.item() - want.item()) < 1e-12 and torch.allclose(A2.grad, A.grad, atol=1e-14) and torch.allclose(B2.grad, B.grad, atol=1e-14)
    cf = oc.clip_loss_closed_form(a.numpy(), b.numpy(), scale)
    assert abs(cf.loss - want.item()) < 1e-12 and abs(cf.dscale - s.grad.item()) < 1e-12
    assert np.allclose(cf.dA, A.grad.numpy(), atol=1e-13) and np.allclose(cf.dB, B.grad.numpy(), atol=1e-13)
    # the row-panel sample of the CPU baseline is the reference's own rows: mean over the first m rows of both matrices
    m = n // 2
    za, zb = (scale * a) @ b.T, (scale * b) @ a.T
    lab = torch.arange(m)
    panel = (torch.nn.functional.cross_entropy(za[:m], lab) + torch.nn.functional.cross_entropy(zb[:m], lab)) / 2
    assert abs(oc.clip_loss_port_panel(a, b, m, scale).item() - panel.item()) < 1e-12


def test_epilogue_closed_forms_match_the_vendored_reference():
    import torch
    _, be = _ref_or_skip()
    x = torch.randn(17, 24, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    assert np.allclose(oc.normalize_closed_form(x.numpy()), be.Normalize(dim=-1)(x).numpy(), atol=1e-15)
    sc = be.LearnableLogitScaling(learnable=True)
    y, s = oc.logit_scaling_closed_form(x.numpy(), float(sc.log_logit_scale))
    assert np.allclose(y, sc(x).detach().numpy(), rtol=1e-6)


@pytest.mark.parametrize("world", [1, 2, 4])
def test_all_ranks_closed_form_equals_the_per_rank_closed_form(world):
    rng = np.random.default_rng(world)
    N, d = 24, 16
    A = rng.standard_normal((N, d)); A /= np.linalg.norm(A, axis=1, keepdims=True)
    B = rng.standard_normal((N, d)); B /= np.linalg.norm(B, axis=1, keepdims=True)
    g = 1.0 + 0.25 * np.arange(world)
    n = N // world
    for ll in (False, True):
        for gwg in (False, True):
            allr = oc.clip_all_ranks_closed_form(A, B, 7.0, world_size=world, local_loss=ll, gather_with_grad=gwg, grad_outputs=g)
            for r in range(world):
                ref = oc.clip_loss_closed_form(A, B, 7.0, rank=r, world_size=world, local_loss=ll, gather_with_grad=gwg, grad_outputs=g)
                assert abs(allr["loss"][r].item() - ref.loss) < 1e-12 and abs(allr["dscale"][r].item() - ref.dscale) < 1e-12
                assert np.allclose(allr["dA"][r * n:(r + 1) * n].numpy(), ref.dA, atol=1e-13)
                assert np.allclose(allr["dB"][r * n:(r + 1) * n].numpy(), ref.dB, atol=1e-13)
