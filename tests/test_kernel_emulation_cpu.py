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
    assert flag.item() == 1 and np.isnan(loss.item())    # a flushed sum raises the hazard flag and the value becomes NaN
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
@pytest.mark.parametrize("d", [1024, 72, 1152, 1800, 2560])     # whole chunks, one ragged chunk, 5 / 8 chunks, long rows (re-read)
def test_l2norm_scale_and_split(vec, dtype, d):
    rng = np.random.default_rng(6)
    rows = 37
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


# ---------------------------------------------------------------------------------------------------
# csrc/clip_kernels.cu: the tensor-core kernels under the functional TMA / mbarrier / tcgen05 / TMEM
# stand-ins of tests/emu/ptx_emu.h.  FWD / DZ / MAX / GEMM were validated on hardware: they calibrate the
# emulation; every epilogue variant has meanwhile run on B200 as well (round 2).
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def tc(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_bf16.h")):
        pytest.skip("g++ or the CUDA headers are not available")
    out = str(tmp_path_factory.mktemp("emu") / "libclip_emu.so")
    p = subprocess.run([gxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-w", "-I" + CUDA_INC,
                        "-I" + os.path.join(ROOT, "tests", "emu"), "-o", out, os.path.join(ROOT, "tests", "emu", "clip_kernels_emu.cpp")],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    lib = C.CDLL(out)
    lib.emu_s_scratch_floats.restype = C.c_size_t
    return lib


def _unit(rng, n, d, scale=1.0):
    x = torch.nn.functional.normalize(torch.from_numpy(rng.standard_normal((n, d))).float(), dim=-1) * scale
    x = x.to(torch.bfloat16)
    return x, x.double().numpy()


SHAPES = [(300, 300, 72, 0), (256, 512, 128, 256), (130, 700, 64, 400)]       # (n, N, d, row offset): ragged tiles, K tails, panels


@pytest.mark.parametrize("n,N,d,off,sms", [(*sh, sms) for sh in SHAPES for sms in (3, 1)] + [(640, 512, 64, 0, 3)])   # + 3 row chunks
def test_emulated_forward_sums(tc, n, N, d, off, sms):
    tc.emu_set_sms(sms)
    rng = np.random.default_rng(n + N)
    A, Av = _unit(rng, n, d)
    B, Bv = _unit(rng, N, d, 1 / 0.07)
    scale = torch.ones(1); stats = torch.zeros(4)
    stats[0], stats[1] = float((Av ** 2).sum(-1).max()), float((Bv ** 2).sum(-1).max())
    rowsum = torch.empty(n); colsum = torch.empty(N)
    scratch = torch.zeros(tc.emu_s_scratch_floats(n, N))
    tc.emu_fwd_sums(_p(A), _p(B), n, N, d, _p(scale), _p(stats), _p(rowsum), _p(colsum), _p(scratch), None, 0)
    E = np.exp2(LOG2E * (Av @ Bv.T))
    assert np.allclose(rowsum.numpy(), E.sum(1), rtol=1e-5) and np.allclose(colsum.numpy(), E.sum(0), rtol=1e-5)
    # clip_s_kernel<FWD_E>: the same sums bit for bit, plus the exponentials as a bf16 panel clipped to n x N
    ld = (N + 63) // 64 * 64
    kept = torch.full((n + 5, ld), 7.0, dtype=torch.bfloat16)
    rs2 = torch.empty(n); cs2 = torch.empty(N)
    scratch.zero_()
    tc.emu_fwd_sums(_p(A), _p(B), n, N, d, _p(scale), _p(stats), _p(rs2), _p(cs2), _p(scratch), _p(kept), ld)
    assert torch.equal(rs2, rowsum) and torch.equal(cs2, colsum)
    assert torch.all(kept[n:] == 7.0) and torch.all(kept[:, (N + 7) // 8 * 8:] == 7.0)     # rows clipped exactly, columns in 16-byte chunks
    assert np.abs(kept[:n, :N].float().numpy() / E - 1).max() < 2 ** -8 + 1e-5


def test_emulated_forward_uses_exact_maximum_when_the_norm_bound_is_loose(tc):
    """Unnormalised features: the max pass (EPI_MAX) supplies G = max(0, x_max - 100)."""
    tc.emu_set_sms(3)
    rng = np.random.default_rng(9)
    n, d = 256, 64
    A, Av = _t(rng.standard_normal((n, d)), torch.bfloat16)
    B, Bv = _t(rng.standard_normal((n, d)), torch.bfloat16)
    scale = torch.tensor([8.0]); stats = torch.zeros(4)
    stats[0], stats[1] = float((Av ** 2).sum(-1).max()), float((Bv ** 2).sum(-1).max())
    rowsum = torch.empty(n); colsum = torch.empty(n)
    scratch = torch.zeros(tc.emu_s_scratch_floats(n, n))
    tc.emu_fwd_sums(_p(A), _p(B), n, n, d, _p(scale), _p(stats), _p(rowsum), _p(colsum), _p(scratch), None, 0)
    X = 8.0 * LOG2E * (Av @ Bv.T)
    assert stats[3].item() == 1.0 and np.isclose(stats[2].item(), X.max(), rtol=1e-5)
    G = max(0.0, X.max() - 100.0)
    assert G > 0
    assert np.allclose(rowsum.numpy(), np.exp2(X - G).sum(1), rtol=2e-4, atol=1e-30)


@pytest.mark.parametrize("rows,N,d,grow0", [(300, 300, 72, 0), (128, 512, 64, 256), (130, 700, 128, 400)])
@pytest.mark.parametrize("variant,sms", [("dz", 3), ("dz", 1), ("dz_l2", 2), ("siglip", 3), ("siglip", 5)])
def test_emulated_panel_kernels(tc, rows, N, d, grow0, variant, sms):
    tc.emu_set_sms(sms)
    rng = np.random.default_rng(rows + N + d)
    A, Av = _unit(rng, rows, d)
    B, Bv = _unit(rng, N, d, 1 / 0.07)
    scale = torch.ones(1); stats = torch.tensor([1.0, 205.0, 0.0, 0.0])
    wr = torch.from_numpy((1e-3 * rng.random(rows)).astype(np.float32)); wc = torch.from_numpy((1e-3 * rng.random(N)).astype(np.float32))
    dg = torch.from_numpy((0.1 * rng.random(rows)).astype(np.float32))
    ldw = (N + 63) // 64 * 64
    Wz = torch.full((rows, ldw), 7.0, dtype=torch.bfloat16)
    Z = Av @ Bv.T
    idx = np.arange(rows)
    if variant == "siglip":
        bias = torch.tensor([-3.0])
        sig = torch.empty(rows); scr = torch.zeros(tc.emu_s_scratch_floats(rows, N))
        tc.emu_siglip_dz(_p(A), _p(B), rows, N, d, grow0, _p(scale), _p(bias), _p(wr), _p(dg), _p(Wz), ldw, _p(sig), _p(scr))
        assert np.allclose(sig.numpy(), (1 / (1 + np.exp(-(Z - 3.0)))).sum(1), rtol=1e-5)        # row sums of sigma (d logit_bias)
        want = wr.double().numpy()[:, None] / (1 + np.exp(-(Z - 3.0)))
    else:
        tc.emu_dz_panel(_p(A), _p(B), rows, N, d, grow0, _p(scale), _p(stats), _p(wr), _p(wc), _p(dg), _p(Wz), ldw, int(variant == "dz_l2"))
        want = np.exp2(LOG2E * Z) * (wr.double().numpy()[:, None] + wc.double().numpy()[None, :])
    want[idx, grow0 + idx] -= dg.double().numpy()
    got = Wz.double().numpy()
    assert np.allclose(got[:, :N], want, rtol=2.0 ** -7, atol=1e-6)            # bf16 storage
    assert np.all(got[:, (N + 7) // 8 * 8:] == 7.0)                            # TMA stores clip whole 16-byte chunks beyond N


@pytest.mark.parametrize("n,N,d", [(300, 300, 72), (256, 512, 128)])
def test_emulated_siglip_forward_and_rowcol_max(tc, n, N, d):
    tc.emu_set_sms(3)
    rng = np.random.default_rng(n * 3 + d)
    A, Av = _unit(rng, n, d)
    B, Bv = _unit(rng, N, d)
    scale = torch.tensor([10.0]); bias = torch.tensor([-10.0])
    rowsum = torch.empty(n)
    scratch = torch.zeros(tc.emu_s_scratch_floats(n, N))
    tc.emu_siglip_fwd(_p(A), _p(B), n, N, d, _p(scale), _p(bias), _p(rowsum), _p(scratch))
    z = 10.0 * (Av @ Bv.T) - 10.0
    sp = np.maximum(z, 0) + np.log1p(np.exp(-np.abs(z)))
    assert np.allclose(rowsum.numpy() * np.log(2.0), sp.sum(1), rtol=2e-5)
    # extreme logits stay finite (2^x alone would overflow)
    big = torch.tensor([4000.0])
    tc.emu_siglip_fwd(_p(A), _p(B), n, N, d, _p(big), None, _p(rowsum), _p(scratch))
    zb = 4000.0 * (Av @ Bv.T)
    assert np.all(np.isfinite(rowsum.numpy())) and np.allclose(rowsum.numpy() * np.log(2.0), (np.maximum(zb, 0) + np.log1p(np.exp(-np.abs(zb)))).sum(1), rtol=1e-4)
    rowmax = torch.empty(n); colmax = torch.empty(N)
    tc.emu_rowcol_max(_p(A), _p(B), n, N, d, _p(scale), _p(rowmax), _p(colmax), _p(scratch))
    X = 10.0 * LOG2E * (Av @ Bv.T)
    assert np.allclose(rowmax.numpy(), X.max(1), rtol=1e-5, atol=1e-5) and np.allclose(colmax.numpy(), X.max(0), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("N,d", [(300, 72), (512, 64)])
def test_emulated_retrieval_ranks(tc, N, d):
    tc.emu_set_sms(3)
    rng = np.random.default_rng(N + d)
    S, Sv = _unit(rng, N, d)
    M = torch.nn.functional.normalize(S.float() + 0.8 * torch.from_numpy(rng.standard_normal((N, d))).float(), dim=-1).to(torch.bfloat16)
    Mv = M.double().numpy()
    Z = Sv @ Mv.T
    label = torch.from_numpy(np.diag(Z).astype(np.float32))
    r1 = torch.empty(N); r2 = torch.empty(N)
    scratch = torch.zeros(tc.emu_s_scratch_floats(N, N))
    tc.emu_retrieval_ranks(_p(S), _p(M), N, d, _p(label), _p(r1), _p(r2), _p(scratch))
    Zf = (S.float() @ M.float().T).double().numpy()          # the kernel compares fp32-accumulated products with the fp32 label dots
    off = ~np.eye(N, dtype=bool)
    d32 = label.double().numpy()
    want1 = ((Z > d32[:, None]) & off).sum(1); want2 = ((Z > d32[None, :]) & off).sum(0)
    assert np.abs(r1.numpy() - want1).max() <= 1 and np.abs(r2.numpy() - want2).max() <= 1      # exact up to fp32-rounding ties
    assert (r1.numpy() == want1).mean() > 0.99 and (r2.numpy() == want2).mean() > 0.99


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,Nc,K", [(300, 264, 200), (128, 256, 64)])
def test_emulated_gemm_all_layouts_and_epilogues(tc, a_mn, b_mn, M, Nc, K):
    tc.emu_set_sms(1 + (M + a_mn + 2 * b_mn) % 4)          # 1..4 persistent CTAs
    rng = np.random.default_rng(M + Nc + K + 2 * a_mn + b_mn)
    A, Av = _t(rng.standard_normal((K, M) if a_mn else (M, K)), torch.bfloat16)
    B, Bv = _t(rng.standard_normal((K, Nc) if b_mn else (Nc, K)), torch.bfloat16)
    ref = (Av.T if a_mn else Av) @ (Bv if b_mn else Bv.T)
    acc_in = torch.from_numpy(rng.standard_normal((M, Nc)).astype(np.float32))
    rs = torch.from_numpy(rng.random(M).astype(np.float32) + 0.5)
    dot, dotv = _t(rng.standard_normal((M, Nc)), torch.bfloat16)
    acc_out = torch.empty(M, Nc); out = torch.empty(M, Nc, dtype=torch.bfloat16)
    ldd = (M + 127) // 128 * 128
    slabs = 2 * ((Nc + 255) // 256)
    rd = torch.zeros(slabs, ldd)
    tc.emu_gemm(_p(A), A.stride(0), a_mn, _p(B), B.stride(0), b_mn, M, Nc, K, _p(acc_in), _p(acc_out), _p(out), Nc, _p(rs), _p(dot), Nc, _p(rd))
    val = ref + acc_in.double().numpy()
    assert np.allclose(acc_out.numpy(), val * rs.double().numpy()[:, None], rtol=1e-5, atol=1e-4)
    assert np.allclose(out.double().numpy(), val * rs.double().numpy()[:, None], rtol=2.0 ** -7, atol=1e-2)
    assert np.allclose(rd.double().numpy()[:, :M].sum(0), (val * dotv).sum(1), rtol=1e-4, atol=1e-3)      # row dots of the unscaled value
