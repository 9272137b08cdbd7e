"""GPU parity of the individual C-ABI kernels against torch fp32/fp64 math on the same inputs.

These go through ctypes -> liboneprot_clip.so (no torch op computes any checked value)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOG2E = 1.4426950408889634


def _k():
    from oneprot_b200 import kernels
    return kernels


def _rand_bf16(r, c, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (scale * torch.randn(r, c, generator=g)).to(torch.bfloat16).cuda()


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,Nc,K", [(128, 256, 64), (256, 512, 256), (300, 264, 200), (1000, 1024, 1000)])
def test_gemm_bf16_all_layouts(a_mn, b_mn, M, Nc, K):
    k = _k()
    # row pitches must be multiples of 8 elements (16 B) for TMA: pad, then view
    r8 = lambda v: (v + 7) // 8 * 8
    A = _rand_bf16(K, r8(M), 1)[:, :M] if a_mn else _rand_bf16(M, r8(K), 1)[:, :K]
    B = _rand_bf16(K, r8(Nc), 2)[:, :Nc] if b_mn else _rand_bf16(Nc, r8(K), 2)[:, :K]
    opA = A.float().T if a_mn else A.float()
    opB = B.float() if b_mn else B.float().T
    ref = (opA.double() @ opB.double())
    out = torch.empty(M, Nc, dtype=torch.bfloat16, device="cuda")
    acc = torch.empty(M, Nc, dtype=torch.float32, device="cuda")
    k.gemm_bf16(A, a_mn, B, b_mn, M, Nc, K, acc_out=acc, out=out)
    torch.cuda.synchronize()
    err = (acc.double() - ref).abs().max().item()
    tol = 1e-4 * math.sqrt(K) + 1e-5
    assert err < tol, f"fp32 accumulate mismatch {err} (tol {tol})"
    assert (out.double() - ref).abs().max().item() < 0.01 * ref.abs().max().item() + 1e-2
    # accumulate-into path (beta = 1)
    acc2 = torch.empty_like(acc)
    k.gemm_bf16(A, a_mn, B, b_mn, M, Nc, K, acc_in=acc, acc_out=acc2)
    torch.cuda.synchronize()
    assert (acc2.double() - 2 * ref).abs().max().item() < 2 * tol


def _fwd_reference(A, B, s):
    Z = s * (A.double() @ B.double().T)
    return Z


@pytest.mark.parametrize("n,N,d,off", [(128, 256, 64, 0), (25, 25, 64, 0), (256, 256, 512, 0), (1000, 1000, 1024, 0),
                                       (300, 900, 128, 300), (2048, 4096, 1024, 2048)])
def test_fwd_sums_match_fp64(n, N, d, off):
    k = _k()
    B_all = _rand_bf16(N, d, 11, 1.0 / math.sqrt(d))
    A = (_rand_bf16(n, d, 12, 1.0 / math.sqrt(d)).float() + 0.5 * B_all[off:off + n].float()).to(torch.bfloat16)
    s = 14.2857
    scale = torch.tensor([s], dtype=torch.float32, device="cuda")
    stats = torch.zeros(4, dtype=torch.float32, device="cuda")
    diag = torch.empty(n, dtype=torch.float32, device="cuda")
    rowsum = torch.empty(n, dtype=torch.float32, device="cuda")
    colsum = torch.empty(N, dtype=torch.float32, device="cuda")
    k.rowstats(A, B_all, off, diag, stats)
    k.fwd_sums(A, B_all, scale, stats, rowsum, colsum)
    torch.cuda.synchronize()
    Z = _fwd_reference(A, B_all, s)
    dref = (A.double() * B_all[off:off + n].double()).sum(-1)
    assert torch.allclose(diag.double(), dref, rtol=1e-5, atol=1e-6)
    assert abs(stats[0].item() - (A.double() ** 2).sum(-1).max().item()) < 1e-4 * stats[0].item()
    assert abs(stats[1].item() - (B_all.double() ** 2).sum(-1).max().item()) < 1e-4 * stats[1].item()
    U = abs(s) * LOG2E * math.sqrt(stats[0].item() * stats[1].item())
    G = max(0.0, U - 100.0)
    # row / column log-sum-exp in natural units
    row_lse = (G + torch.log2(rowsum.double())) / LOG2E
    col_lse = (G + torch.log2(colsum.double())) / LOG2E
    assert torch.allclose(row_lse, torch.logsumexp(Z, dim=1), rtol=0, atol=2e-4)
    assert torch.allclose(col_lse, torch.logsumexp(Z, dim=0), rtol=0, atol=2e-4)


@pytest.mark.parametrize("rows,N,d,grow0", [(128, 256, 64, 0), (25, 25, 64, 0), (300, 900, 128, 300), (1000, 1000, 1024, 0)])
def test_dz_panel_matches_fp64(rows, N, d, grow0):
    k = _k()
    B_all = _rand_bf16(N, d, 21, 1.0 / math.sqrt(d))
    A = _rand_bf16(rows, d, 22, 1.0 / math.sqrt(d))
    s = 10.0
    scale = torch.tensor([s], dtype=torch.float32, device="cuda")
    stats = torch.zeros(4, dtype=torch.float32, device="cuda")
    diag = torch.empty(rows, dtype=torch.float32, device="cuda")
    k.rowstats(A, B_all, grow0, diag, stats)
    g = torch.Generator(device="cpu").manual_seed(5)
    wr = torch.rand(rows, generator=g).cuda()
    wc = torch.rand(N, generator=g).cuda()
    dg = torch.rand(rows, generator=g).cuda()
    ldw = ((N + 63) // 64) * 64
    Wz = torch.zeros(rows, ldw, dtype=torch.bfloat16, device="cuda")
    k.dz_panel(A, B_all, grow0, scale, stats, wr, wc, dg, Wz)
    torch.cuda.synchronize()
    X = s * LOG2E * (A.double() @ B_all.double().T)
    E = torch.exp2(X)   # G = 0 for these magnitudes
    ref = E * (wr.double()[:, None] + wc.double()[None, :])
    idx = torch.arange(rows, device="cuda")
    ref[idx, grow0 + idx] -= dg.double()
    got = Wz[:, :N].double()
    rel = (got - ref).abs().max().item() / ref.abs().max().item()
    assert rel < 1e-2, rel     # bf16 storage: 2^-9 relative


def test_l2norm_fwd_bwd():
    k = _k()
    for dtype in (torch.bfloat16, torch.float32):
        x = (3 * torch.randn(37, 1024, generator=torch.Generator().manual_seed(3))).to(dtype).cuda()
        x[5] = 0
        gy = torch.randn(37, 1024, generator=torch.Generator().manual_seed(4)).to(dtype).cuda()
        sc = torch.tensor([14.2857], dtype=torch.float32, device="cuda")
        y = torch.empty_like(x); gx = torch.empty_like(x)
        inv = torch.empty(37, dtype=torch.float32, device="cuda")
        dsp = torch.empty(37, dtype=torch.float32, device="cuda")
        k.l2norm_scale_fwd(x, y, inv, sc)
        k.l2norm_scale_bwd(x, gy, inv, gx, dsp, sc)
        torch.cuda.synchronize()
        X = x.double().requires_grad_(True)
        S = sc.double().requires_grad_(True)
        Y = S * torch.nn.functional.normalize(X, dim=-1, p=2)
        Y.backward(gy.double())
        tol = 2e-2 if dtype == torch.bfloat16 else 1e-5
        assert torch.allclose(y.double(), Y.detach(), rtol=tol, atol=tol)
        assert torch.allclose(gx.double(), X.grad, rtol=tol, atol=tol * 10)
        assert abs(dsp.double().sum().item() - S.grad.item()) < 1e-2 * abs(S.grad.item()) + 1e-3
