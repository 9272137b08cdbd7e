"""CPU: the SOURCE of the projection-head CUDA kernels (oneprot_b200/csrc/head_kernels.cu) compiled for
the host through the SIMT emulation of tests/emu/cuda_emu.h (one OS thread per CUDA thread, barriers for
__syncthreads and the warp shuffles), launched with the grid shapes of the CUDA host code and compared
with the numpy float64 oracle (oracle/head_oracle.py).  This checks indexing, reductions and arithmetic
of the kernels themselves where no GPU exists; the tensor-core kernels cannot be emulated this way."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import head_oracle as ho

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_bf16.h")):
        pytest.skip("g++ or the CUDA headers are not available")
    out = str(tmp_path_factory.mktemp("emu") / "libhead_emu.so")
    p = subprocess.run([gxx, "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-w", "-I" + CUDA_INC, "-o", out,
                        os.path.join(ROOT, "tests", "emu", "head_kernels_emu.cpp")], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    return C.CDLL(out)


def _t(x, dtype):
    """numpy float64 -> (torch tensor in dtype, its float64 values)"""
    t = torch.from_numpy(np.ascontiguousarray(x)).to(dtype).contiguous()
    return t, t.double().numpy()


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _tol(dtype):
    return (2e-5, 2e-5) if dtype == torch.float32 else (2e-2, 2e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,d", [(70, 1280), (9, 64), (33, 1152), (5, 2048), (11, 2560)])
def test_layernorm_kernels(emu, rows, d, dtype):
    rng = np.random.default_rng(rows * 7 + d)
    x, xv = _t(2.0 * rng.standard_normal((rows, d)) + 0.5, dtype)
    g, gv = _t(1.0 + 0.3 * rng.standard_normal(d), dtype)
    b, bv = _t(0.2 * rng.standard_normal(d), dtype)
    gy, gyv = _t(rng.standard_normal((rows, d)), dtype)
    y = torch.empty_like(x); gx = torch.empty_like(x)
    mean = torch.empty(rows); rstd = torch.empty(rows)
    dg = torch.empty(d); db = torch.empty(d)
    f32 = int(dtype == torch.float32)
    emu.emu_layernorm_fwd(_p(x), _p(g), _p(b), _p(y), _p(mean), _p(rstd), rows, d, f32, C.c_float(1e-5))
    scratch = torch.empty(emu.emu_ln_scratch_floats(rows, d))
    emu.emu_layernorm_bwd(_p(x), _p(gy), _p(g), _p(mean), _p(rstd), _p(gx), _p(dg), _p(db), _p(scratch), rows, d, f32)
    yr, cache = ho.layernorm_fwd(xv, gv, bv)
    gxr, dgr, dbr = ho.layernorm_bwd(gyv, gv, cache)
    rt, at = _tol(dtype)
    assert np.allclose(y.double().numpy(), yr, rtol=rt, atol=at)
    assert np.allclose(mean.numpy(), xv.mean(-1), rtol=1e-5, atol=1e-5) and np.allclose(rstd.numpy(), cache[1][:, 0], rtol=1e-4)
    assert np.allclose(gx.double().numpy(), gxr, rtol=rt, atol=at * 3)
    assert np.allclose(dg.numpy(), dgr, rtol=1e-4, atol=1e-4) and np.allclose(db.numpy(), dbr, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gelu_kernels(emu, dtype):
    rng = np.random.default_rng(3)
    x, xv = _t(3.0 * rng.standard_normal(40 * 264), dtype)
    gy, gyv = _t(rng.standard_normal(40 * 264), dtype)
    y = torch.empty_like(x); gx = torch.empty_like(x)
    f32 = int(dtype == torch.float32)
    emu.emu_gelu(_p(x), None, _p(y), C.c_size_t(x.numel()), f32)
    emu.emu_gelu(_p(x), _p(gy), _p(gx), C.c_size_t(x.numel()), f32)
    rt, at = _tol(dtype)
    assert np.allclose(y.double().numpy(), ho.gelu_fwd(xv), rtol=rt, atol=at)
    assert np.allclose(gx.double().numpy(), ho.gelu_bwd(gyv, xv), rtol=rt, atol=at)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,L,D,masked", [(5, 19, 1280, True), (3, 8, 72, True), (4, 33, 264, False)])
def test_meanpool_kernels(emu, B, L, D, masked, dtype):
    rng = np.random.default_rng(B + L + D)
    x, xv = _t(rng.standard_normal((B, L, D)), dtype)
    gy, gyv = _t(rng.standard_normal((B, D)), dtype)
    mask = None
    if masked:
        lens = rng.integers(1, L + 1, B)
        lens[0] = 1
        mask = (np.arange(L)[None, :] < lens[:, None]).astype(np.float32)
    m = None if mask is None else torch.from_numpy(mask)
    y = torch.empty(B, D, dtype=dtype); inv = torch.empty(B); gx = torch.empty_like(x)
    f32 = int(dtype == torch.float32)
    emu.emu_meanpool_fwd(_p(x), _p(m), _p(y), _p(inv), B, L, D, f32, 1)
    emu.emu_meanpool_bwd(_p(gy), _p(m), _p(inv), _p(gx), B, L, D, f32)
    rt, at = _tol(dtype)
    assert np.allclose(y.double().numpy(), ho.meanpool_fwd(xv, mask), rtol=rt, atol=at)
    assert np.allclose(gx.double().numpy(), ho.meanpool_bwd(gyv, mask, xv.shape), rtol=rt, atol=at)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,L,D", [(4, 21, 1280), (3, 300, 64)])
def test_attention_pooling_kernels(emu, B, L, D, dtype):
    """token_dot -> softmax_rows -> weighted sum, and the backward chain, as heads._AttnPoolFn issues them."""
    rng = np.random.default_rng(L + D)
    x, xv = _t(rng.standard_normal((B, L, D)), dtype)
    w, wv = _t(3.0 * rng.standard_normal(D) / np.sqrt(D), dtype)
    bias = torch.tensor([0.3])
    gy, gyv = _t(rng.standard_normal((B, D)), dtype)
    lens = rng.integers(1, L + 1, B)
    lens[0], lens[-1] = 1, L
    mask = (np.arange(L)[None, :] < lens[:, None]).astype(np.float32)
    m = torch.from_numpy(mask)
    f32 = int(dtype == torch.float32)
    p = torch.empty(B, L)
    emu.emu_token_dot(_p(x), _p(w), 0, _p(bias), _p(m), _p(p), B, L, D, f32)
    emu.emu_softmax_rows(_p(p), _p(p), B, L)
    y = torch.empty(B, D, dtype=dtype)
    emu.emu_meanpool_fwd(_p(x), _p(p), _p(y), None, B, L, D, f32, 0)
    yr, pr = ho.attnpool_fwd(xv, mask, wv, 0.3)
    rt, at = _tol(dtype)
    assert np.allclose(p.numpy(), pr, rtol=1e-4, atol=1e-6)
    assert np.allclose(y.double().numpy(), yr, rtol=rt, atol=at)
    dp = torch.empty(B, L); ds = torch.empty(B, L); gx = torch.empty_like(x)
    emu.emu_token_dot(_p(x), _p(gy), 1, None, None, _p(dp), B, L, D, f32)
    emu.emu_softmax_rows_bwd(_p(p), _p(dp), _p(ds), B, L)
    emu.emu_attnpool_bwd_x(_p(gy), _p(p), _p(ds), _p(w), _p(gx), B, L, D, f32)
    part = torch.empty(B, D, dtype=dtype)
    emu.emu_meanpool_fwd(_p(x), _p(ds), _p(part), None, B, L, D, f32, 0)
    gxr, gwr, gbr = ho.attnpool_bwd(gyv, xv, wv, pr)
    assert np.allclose(gx.double().numpy(), gxr, rtol=rt, atol=at)
    assert np.allclose(part.double().numpy().sum(0), gwr, rtol=rt * 5, atol=at * 5)
    assert abs(ds.double().numpy().sum() - gbr) < 1e-4


# ---------------------------------------------------------------------------------------------------
# csrc/vector_kernels.cuh: every non-tensor-core kernel of the ClipLoss path
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def vec(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_bf16.h")):
        pytest.skip("g++ or the CUDA headers are not available")
    out = str(tmp_path_factory.mktemp("emu") / "libvec_emu.so")
    p = subprocess.run([gxx, "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-w", "-I" + CUDA_INC, "-o", out,
                        os.path.join(ROOT, "tests", "emu", "vector_kernels_emu.cpp")], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    return C.CDLL(out)


LOG2E = 1.4426950408889634


def test_rowstats_and_rowdots(vec):
    rng = np.random.default_rng(1)
    n, N, d, off = 40, 100, 72, 30
    A, Av = _t(rng.standard_normal((n, d)), torch.bfloat16)
    B, Bv = _t(2.0 * rng.standard_normal((N, d)), torch.bfloat16)
    diag = torch.empty(n); stats = torch.zeros(4)
    vec.emu_rowstats(_p(A), _p(B), n, N, d, off, _p(diag), _p(stats))
    assert np.allclose(diag.numpy(), (Av * Bv[off:off + n]).sum(-1), rtol=1e-5, atol=1e-5)
    assert np.isclose(stats[0].item(), (Av * Av).sum(-1).max(), rtol=1e-5) and np.isclose(stats[1].item(), (Bv * Bv).sum(-1).max(), rtol=1e-5)
    out = torch.empty(n)
    vec.emu_rowdot_bf16(_p(A), d, _p(B), d, n, d, _p(out))
    assert np.allclose(out.numpy(), (Av * Bv[:n]).sum(-1), rtol=1e-5, atol=1e-5)
    for dtype in (torch.float32, torch.bfloat16):
        X, Xv = _t(rng.standard_normal((n, d)), dtype)
        Y, Yv = _t(rng.standard_normal((n, d)), dtype)
        vec.emu_rowdot(_p(X), _p(Y), n, d, int(dtype == torch.float32), _p(out))
        assert np.allclose(out.numpy(), (Xv * Yv).sum(-1), rtol=1e-5, atol=1e-5)
    v = torch.from_numpy(rng.standard_normal(5000).astype(np.float32))
    t = torch.empty(1)
    vec.emu_sum_f32(_p(v), 5000, _p(t))
    assert np.isclose(t.item(), v.double().sum().item(), rtol=1e-6, atol=1e-5)


def test_slot_reductions_and_augment(vec):
    rng = np.random.default_rng(2)
    slots, ld, count = 13, 300, 270
    part = torch.from_numpy(rng.standard_normal((slots, ld)).astype(np.float32))
    out = torch.empty(count)
    vec.emu_reduce_slots(_p(part), slots, ld, count, _p(out), 0)
    assert np.allclose(out.numpy(), part.double().numpy()[:, :count].sum(0), rtol=1e-5, atol=1e-5)
    vec.emu_reduce_slots(_p(part), slots, ld, count, _p(out), 1)
    assert np.array_equal(out.numpy(), part.numpy()[:, :count].max(0))
    W, cnt = 5, 8 * 77
    sl, slv = _t(rng.standard_normal((W, cnt)), torch.bfloat16)
    o = torch.empty(cnt, dtype=torch.bfloat16)
    vec.emu_sum_slots_bf16(_p(sl), W, C.c_size_t(cnt), _p(o))
    assert torch.equal(o, torch.from_numpy(slv.sum(0)).to(torch.bfloat16))
    # operand augmentation of the two-reference path: two bf16 limbs of -ref / c, or ones
    rows, d = 19, 64
    x, xv = _t(rng.standard_normal((rows, d)), torch.bfloat16)
    ref = torch.from_numpy((1e5 * rng.standard_normal(rows)).astype(np.float32))
    scale = torch.tensor([3.0])
    aug = torch.empty(rows, d + 8, dtype=torch.bfloat16); refq = torch.empty(rows)
    vec.emu_augment(_p(x), rows, d, _p(ref), _p(scale), _p(aug), _p(refq))
    c = 3.0 * LOG2E
    a = aug.double().numpy()
    assert np.array_equal(a[:, :d], xv) and np.all(a[:, d + 2:] == 0)
    assert np.allclose(-c * (a[:, d] + a[:, d + 1]), refq.numpy(), rtol=1e-6)
    assert np.all(np.abs(refq.numpy() - ref.numpy()) <= 2.0 ** -15 * np.abs(ref.numpy()) + 1e-3)      # two limbs: ~2^-17 relative
    vec.emu_augment(_p(x), rows, d, None, _p(scale), _p(aug), None)
    assert np.all(aug.double().numpy()[:, d:d + 2] == 1.0)


@pytest.mark.parametrize("mode,with_refs", [(0, False), (1, False), (0, True)])
def test_loss_finalize_and_bwd_weights(vec, mode, with_refs):
    rng = np.random.default_rng(3 + mode)
    N, n, off, s = 700, 350, 350, 2.5
    rowsum = torch.from_numpy(np.exp(rng.standard_normal(N)).astype(np.float32))
    colsum = torch.from_numpy(np.exp(rng.standard_normal(N)).astype(np.float32))
    diag = torch.from_numpy(rng.standard_normal(N).astype(np.float32))
    scale = torch.tensor([s]); stats = torch.tensor([4.0, 9.0, 0.0, 0.0])          # U = c * 6 < 100  =>  G = 0
    rr = torch.from_numpy((50 * rng.standard_normal(N)).astype(np.float32)) if with_refs else None
    cr = torch.from_numpy((50 * rng.standard_normal(N)).astype(np.float32)) if with_refs else None
    loss = torch.empty(1); inv_rs = torch.empty(N); inv_cs = torch.empty(N)
    flag = torch.zeros(1, dtype=torch.int32); scratch = torch.zeros(64, dtype=torch.float64)
    counter = scratch[32:].view(torch.int32)
    for _ in range(2):                                   # twice: the kernel leaves its counter ready for the next launch
        vec.emu_loss_finalize(_p(rowsum), _p(colsum), _p(diag), N, n, off, mode, _p(scale), _p(stats), _p(loss), _p(inv_rs), _p(inv_cs),
                              _p(flag), _p(scratch), C.c_void_p(counter.data_ptr()), _p(rr), _p(cr))
        lo, hi = (off, off + n) if mode == 1 else (0, N)
        gr = rr.double().numpy() if with_refs else 0.0
        gc = cr.double().numpy() if with_refs else 0.0
        zd = s * diag.double().numpy()
        rl = (gr + np.log2(rowsum.double().numpy())) / LOG2E - zd
        cl = (gc + np.log2(colsum.double().numpy())) / LOG2E - zd
        want = (rl[lo:hi].sum() + cl[lo:hi].sum()) / (2 * (hi - lo))
        assert np.isclose(loss.item(), want, rtol=1e-5, atol=1e-5)
        assert np.allclose(inv_rs.numpy(), 1 / rowsum.numpy(), rtol=1e-6) and flag.item() == 0
    colsum[5] = 0.0
    vec.emu_loss_finalize(_p(rowsum), _p(colsum), _p(diag), N, n, off, mode, _p(scale), _p(stats), _p(loss), _p(inv_rs), _p(inv_cs),
                          _p(flag), _p(scratch), C.c_void_p(counter.data_ptr()), _p(rr), _p(cr))
    assert flag.item() == 1                              # a flushed sum raises the hazard flag
    # backward weights of both conventions
    world, rank = 2, 1
    gvec = torch.tensor([1.0, 1.5]); wr = torch.empty(n); wc = torch.empty(N); dg = torch.empty(n); sa = torch.empty(n); sb = torch.empty(N)
    vec.emu_bwd_weights(_p(inv_rs), _p(inv_cs), N, n, off, mode, 1, 0, world, rank, _p(gvec), _p(scale), _p(wr), _p(wc), _p(dg), _p(sa), _p(sb), 0)
    if mode == 0:
        coef = s / (2 * N)
        assert np.allclose(wr.numpy(), coef * inv_rs.numpy()[off:off + n], rtol=1e-6) and np.allclose(dg.numpy(), 2 * coef)
        assert np.allclose(sa.numpy(), 2.5) and np.allclose(sb.numpy(), 2.5)          # use_gsum: sum of the upstream gradients
    else:
        assert np.allclose(wr.numpy(), s * 1.5 / (2 * n) * inv_rs.numpy()[off:off + n], rtol=1e-6)
        owner_g = np.repeat([1.0, 1.5], n)
        assert np.allclose(wc.numpy(), s * owner_g / (2 * n) * inv_cs.numpy(), rtol=1e-6, atol=1e-30, equal_nan=True)


def test_siglip_finalize(vec):
    rng = np.random.default_rng(5)
    n = 3000
    rowsum = torch.from_numpy((10 + rng.random(n)).astype(np.float32)); diag = torch.from_numpy(rng.standard_normal(n).astype(np.float32))
    scale = torch.tensor([10.0]); bias = torch.tensor([-10.0]); loss = torch.empty(1)
    for b in (bias, None):
        vec.emu_siglip_finalize(_p(rowsum), _p(diag), n, _p(scale), _p(b), _p(loss))
        want = (np.log(2.0) * rowsum.double().numpy().sum() - (10.0 * diag.double().numpy() + (0.0 if b is None else -10.0)).sum()) / n
        assert np.isclose(loss.item(), want, rtol=1e-6)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_l2norm_scale_and_split(vec, dtype):
    rng = np.random.default_rng(6)
    rows, d = 37, 1024
    x, xv = _t(3 * rng.standard_normal((rows, d)), dtype)
    x[5] = 0
    xv[5] = 0
    gy, gyv = _t(rng.standard_normal((rows, d)), dtype)
    sc = torch.tensor([14.2857]); y = torch.empty_like(x); gx = torch.empty_like(x); inv = torch.empty(rows); dsp = torch.empty(rows)
    f32 = int(dtype == torch.float32)
    vec.emu_l2norm_fwd(_p(x), _p(y), _p(inv), rows, d, f32, _p(sc), C.c_float(1e-12))
    vec.emu_l2norm_bwd(_p(x), _p(gy), _p(inv), _p(gx), _p(dsp), rows, d, f32, _p(sc), C.c_float(1e-12))
    from oracle import clip_oracle as oc
    yr = 14.2857 * oc.normalize_closed_form(xv)
    gxr = 14.2857 * oc.normalize_backward_closed_form(xv, gyv)
    rt, at = _tol(dtype)
    assert np.allclose(y.double().numpy(), yr, rtol=rt, atol=at) and np.allclose(gx.double().numpy(), gxr, rtol=rt, atol=at * 10)
    assert np.allclose(dsp.numpy(), (oc.normalize_closed_form(xv) * gyv).sum(-1), rtol=1e-3, atol=1e-3)
    vec.emu_scale_rows(_p(x), _p(y), rows, d, f32, _p(sc))
    assert np.allclose(y.double().numpy(), 14.2857 * xv, rtol=rt, atol=at)
    if dtype == torch.float32:
        out = torch.empty(rows, 3 * d, dtype=torch.bfloat16)
        for side in (0, 1):
            vec.emu_split_fp32(_p(x), _p(out), rows, d, side, 3)
            o = out.double().numpy()
            h, m = (o[:, :d], o[:, 2 * d:]) if side == 0 else (o[:, :d], o[:, d:2 * d])
            assert np.allclose(h + m, xv, rtol=2.0 ** -15, atol=1e-30)       # two limbs reproduce fp32 to ~2^-16
            assert np.array_equal(o[:, d:2 * d] if side == 0 else o[:, 2 * d:], h)
