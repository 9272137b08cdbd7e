"""CPU: the stored-exponentials backward (ClipLoss(keep_exp=True)) - host logic over the float64 stand-ins and
over the library's own kernel source under the CPU emulation (clip_s_kernel<FWD_E>, dz_from_exp_kernel)."""
import numpy as np
import pytest
import torch

from oracle import clip_oracle as oc
from tests import fake_kernels
from tests.helpers import bf16_from_bits, cosine, load_golden, rel_err


def _provider(name):
    if name == "fake":
        return fake_kernels
    from tests import emu_kernels
    if not emu_kernels.available():
        pytest.skip("g++ or the CUDA headers are not available")
    emu_kernels.prebuild()
    return emu_kernels


@pytest.fixture(params=["fake", "emu"])
def prov(request, monkeypatch):
    from oneprot_b200 import clip_loss
    K = _provider(request.param)
    monkeypatch.setattr(clip_loss, "_KERNELS", K)
    clip_loss._SCALE_CACHE.clear()
    return clip_loss, K


@pytest.mark.parametrize("name", ["clip_single_n25_d64_train.npz", "clip_single_n100_d72_scale.npz", "clip_single_n96_d128_uncorr.npz"])
def test_keep_exp_matches_reference_golden_and_skips_the_recompute(prov, name):
    cl, K = prov
    g = load_golden(name)
    is_t = bool(g["scale_is_tensor"])
    res = {}
    for keep in (False, True):
        A = bf16_from_bits(g["A_bf16"]).requires_grad_(True)
        B = bf16_from_bits(g["B_bf16"]).requires_grad_(True)
        ls = torch.tensor(float(g["scale"]), requires_grad=True) if is_t else float(g["scale"])
        m = cl.ClipLoss(loss_dtype=torch.float32, keep_exp=keep, panel_bytes=128 * 128 * 2)
        del K.CALLS[:]
        loss = m(A, B, ls)
        loss.backward()
        calls = list(K.CALLS)
        if keep:
            assert "fwd_sums_keep" in calls and "dz_from_exp" in calls and "dz_panel" not in calls and "fwd_sums" not in calls
            # one panel: one dB GEMM and one dA GEMM
            assert calls.count("gemm") == 2
        else:
            assert "dz_panel" in calls and "dz_from_exp" not in calls
        assert rel_err(loss.item(), g["loss_f64"]) < 1e-5
        assert cosine(A.grad.float().numpy(), g["dA_f64"]) > 0.9999
        assert cosine(B.grad.float().numpy(), g["dB_f64"]) > 0.9999
        if is_t:
            assert rel_err(ls.grad.item(), g["dscale_f64"]) < 2e-2
        res[keep] = (loss.item(), A.grad.float().numpy(), B.grad.float().numpy())
    assert res[True][0] == res[False][0]                      # the forward sums are the same arithmetic
    for k in (1, 2):                                          # one more bf16 rounding of e_ij, nothing else
        assert cosine(res[True][k], res[False][k]) > 0.99999
        assert abs(np.linalg.norm(res[True][k]) / np.linalg.norm(res[False][k]) - 1) < 2e-3


def test_second_backward_recomputes(prov):
    """E is overwritten by the first backward; retain_graph + a second backward must not reuse it."""
    cl, K = prov
    g = load_golden("clip_single_n25_d64_train.npz")
    A = bf16_from_bits(g["A_bf16"]).requires_grad_(True)
    B = bf16_from_bits(g["B_bf16"]).requires_grad_(True)
    loss = cl.ClipLoss(loss_dtype=torch.float32, keep_exp=True)(A, B)
    loss.backward(retain_graph=True)
    first = A.grad.float().clone()
    A.grad = None
    del K.CALLS[:]
    loss.backward()
    assert "dz_panel" in K.CALLS and "dz_from_exp" not in K.CALLS
    assert cosine(A.grad.float().numpy(), first.numpy()) > 0.99999
    assert cosine(A.grad.float().numpy(), g["dA_f64"]) > 0.9999


def test_keep_exp_is_skipped_where_it_does_not_apply(prov):
    cl, K = prov
    gen = torch.Generator().manual_seed(5)
    a = torch.nn.functional.normalize(torch.randn(40, 32, generator=gen), dim=-1)
    b = torch.nn.functional.normalize(torch.randn(40, 32, generator=gen), dim=-1)
    # no gradient needed / panel larger than keep_bytes / fp32 inputs (limb split): the plain forward runs
    for kw, A, B in (({}, a.to(torch.bfloat16), b.to(torch.bfloat16)),
                     ({"keep_bytes": 1024}, a.to(torch.bfloat16).requires_grad_(True), b.to(torch.bfloat16)),
                     ({}, a.clone().requires_grad_(True), b.clone())):
        del K.CALLS[:]
        loss = cl.ClipLoss(keep_exp=True, loss_dtype=torch.float32, **kw)(A, B, 5.0)
        if A.requires_grad:
            loss.backward()
        assert "fwd_sums_keep" not in K.CALLS and "dz_from_exp" not in K.CALLS
        ref = oc.clip_loss_closed_form(A.detach().double().numpy(), B.detach().double().numpy(), 5.0)
        assert rel_err(loss.item(), ref.loss) < 1e-5


@pytest.mark.parametrize("n,N,grow0", [(5, 13, 0), (40, 300, 17), (33, 2056, 2000)])
def test_dz_from_exp_kernel_source_matches_float64(n, N, grow0):
    """dz_from_exp_kernel itself (CPU SIMT emulation): ragged right edge, diagonal offsets, row pitch > N."""
    K = _provider("emu")
    gen = torch.Generator().manual_seed(n + N)
    ld = (N + 63) // 64 * 64
    E = torch.rand(n + 3, ld, generator=gen).to(torch.bfloat16)
    E0 = E.clone()
    wr, dg = torch.rand(n, generator=gen), torch.rand(n, generator=gen)
    wc = torch.rand(N, generator=gen)
    K.dz_from_exp(E, n, N, grow0, wr, wc, dg)
    want = E0.clone()
    fake_kernels.dz_from_exp(want, n, N, grow0, wr, wc, dg)
    assert torch.equal(E[n:], E0[n:]) and torch.equal(E[:, N:], E0[:, N:])          # nothing outside rows x N is touched
    assert torch.equal(E[:n, :N], want[:n, :N])                                      # bf16 round-to-nearest of the same fp32 value


@pytest.mark.parametrize("n,N,d,off", [(25, 25, 64, 0), (130, 700, 72, 400), (256, 512, 128, 256)])
def test_siglip_keeping_forward_kernel_source_matches_float64(n, N, d, off):
    """clip_s_kernel<SFWD_K> (CPU emulation): softplus row sums as SFWD, S = sigma(z) - [i == j] as a bf16 panel clipped to
    n x N, row sums of sigma."""
    K = _provider("emu")
    gen = torch.Generator().manual_seed(n + N)
    A = torch.nn.functional.normalize(torch.randn(n, d, generator=gen), dim=-1).to(torch.bfloat16)
    B = torch.nn.functional.normalize(torch.randn(N, d, generator=gen), dim=-1).to(torch.bfloat16)
    scale, bias = torch.tensor([9.0]), torch.tensor([-4.0])
    ld = (N + 63) // 64 * 64
    out = {}
    for name, P in (("emu", K), ("fake", fake_kernels)):
        S = torch.full((n + 3, ld), 7.0, dtype=torch.bfloat16)
        rs, sig = torch.empty(n), torch.empty(n)
        P.siglip_fwd_keep(A, B, off, scale, bias, rs, S, sig_rowsum=sig)
        out[name] = (S, rs, sig)
    S, rs, sig = out["emu"]
    S0, rs0, sig0 = out["fake"]
    rs_plain = torch.empty(n)
    K.siglip_fwd(A, B, scale, bias, rs_plain)
    assert torch.equal(rs, rs_plain)                                                  # the softplus sums are SFWD's
    assert np.allclose(rs.numpy(), rs0.numpy(), rtol=1e-5) and np.allclose(sig.numpy(), sig0.numpy(), rtol=1e-5)
    assert torch.all(S[n:] == 7.0) and torch.all(S[:, (N + 7) // 8 * 8:] == 7.0)   # clipped stores (whole 16-byte chunks: padding columns are scratch)
    assert float((S[:n, :N].float() - S0[:n, :N].float()).abs().max()) <= 2 ** -8     # one bf16 ulp of values in [-1, 1]
    idx = torch.arange(n)
    assert torch.all(S[idx, off + idx].float() < 0) and torch.all(S[:n, :N].float() <= 1)
