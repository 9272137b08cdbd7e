"""Kernel provider backed by the CPU EMULATION of the library's own CUDA kernel source (tests/emu):
same interface as oneprot_b200.kernels (and tests/fake_kernels.py), but every call runs the real kernel
bodies of csrc/clip_kernels.cu / vector_kernels.cuh / head_kernels.cu under the SIMT + TMA / mbarrier /
tcgen05 stand-ins, with the parameter set-up of the CUDA host functions.  Injected like the float64
emulation (``clip_loss._KERNELS = emu_kernels``) it lets the CPU tests run the Python host of the product
over its actual kernels.  Test infrastructure only; slow (scalar MMAs), meant for small problems."""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"
MODE_GLOBAL, MODE_LOCAL = 0, 1
CALLS = []
_LIBS = {}


def available():
    return shutil.which("g++") is not None and os.path.exists(os.path.join(CUDA_INC, "cuda_bf16.h"))


def prebuild():
    """Compile all three emulation libraries (call once before spawning ranks)."""
    _tc(); _vec(); _head()


def _stamp():
    """Cache key of a build: newest modification time of the kernel and emulation sources."""
    dirs = [os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "oneprot_b200", "csrc"), os.path.join(ROOT, "include")]
    return str(int(max(os.path.getmtime(os.path.join(d, f)) for d in dirs for f in os.listdir(d))))


def _lib(name):
    if name not in _LIBS:
        cache = os.path.join(tempfile.gettempdir(), "oneprot_emu_cache_" + _stamp())
        os.makedirs(cache, exist_ok=True)
        out = os.path.join(cache, f"lib{name}.so")
        if os.path.exists(out):                    # built by an earlier test / another rank
            _LIBS[name] = C.CDLL(out)
            if name == "clip_kernels_emu":
                _LIBS[name].emu_s_scratch_floats.restype = C.c_size_t
            return _LIBS[name]
        tmp = out + f".{os.getpid()}.tmp"
        cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-w", "-I" + CUDA_INC, "-I" + os.path.join(ROOT, "tests", "emu"),
               "-o", tmp, os.path.join(ROOT, "tests", "emu", name + ".cpp")]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(p.stderr[-3000:])
        os.replace(tmp, out)                       # atomic: concurrent ranks never load a half-written file
        lib = C.CDLL(out)
        if name == "clip_kernels_emu":
            lib.emu_s_scratch_floats.restype = C.c_size_t
        _LIBS[name] = lib
    return _LIBS[name]


def _tc():
    return _lib("clip_kernels_emu")


def _vec():
    return _lib("vector_kernels_emu")


def _head():
    return _lib("head_kernels_emu")


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32(t):
    return int(t.dtype == torch.float32)


class stream_scope:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def panel_row_unit(d):
    return 128


def launch_count():
    return len(CALLS)


def launch_count_reset():
    CALLS.clear()


def current_stream_handle():
    return 0


# ---- ClipLoss path ---------------------------------------------------------------------------------
def rowstats(A, B_all, row_offset, diag, stats):
    CALLS.append("rowstats")
    _vec().emu_rowstats(_p(A), _p(B_all), A.shape[0], B_all.shape[0], A.shape[1], row_offset, _p(diag), _p(stats))


def fwd_scratch_bytes(n, N):
    return 4 * int(_tc().emu_s_scratch_floats(n, N))


def _scratch(n, N):
    return torch.zeros(int(_tc().emu_s_scratch_floats(n, N)), dtype=torch.float32)


def fwd_sums(A, B_all, scale_dev, stats, rowsum, colsum, scratch=None, ag=None, keep=None):
    assert ag is None, "the fused all-gather is not emulated"
    CALLS.append("fwd_sums" if keep is None else "fwd_sums_keep")
    n, d = A.shape
    N = B_all.shape[0]
    s = _scratch(n, N)
    _tc().emu_fwd_sums(_p(A), _p(B_all), n, N, d, _p(scale_dev), _p(stats), _p(rowsum), _p(colsum), _p(s), _p(keep),
                       keep.stride(0) if keep is not None else 0)
    return scratch


def dz_from_exp(E, rows, N, grow0, wr, wc, dg):
    CALLS.append("dz_from_exp")
    _vec().emu_dz_from_exp(_p(E), rows, N, E.stride(0), grow0, _p(wr), _p(wc), _p(dg))


def loss_finalize(rowsum_all, colsum_all, diag_all, n, row_offset, mode, scale_dev, stats, loss_out, inv_rowsum, inv_colsum, flag,
                  row_ref=None, col_ref=None):
    CALLS.append("loss_finalize")
    N = rowsum_all.numel()
    scratch = torch.zeros(64, dtype=torch.float64)
    counter = scratch[32:].view(torch.int32)
    _vec().emu_loss_finalize(_p(rowsum_all), _p(colsum_all), _p(diag_all), N, n, row_offset, mode, _p(scale_dev), _p(stats), _p(loss_out),
                             _p(inv_rowsum), _p(inv_colsum), _p(flag), _p(scratch), C.c_void_p(counter.data_ptr()), _p(row_ref), _p(col_ref))


def bwd_weights(inv_rowsum, inv_colsum, n, row_offset, mode, use_gsum, part, world, rank, gvec, scale_dev, wr, wc, dg, out_scale_a,
                out_scale_b, what=0):
    CALLS.append("bwd_weights")
    _vec().emu_bwd_weights(_p(inv_rowsum), _p(inv_colsum), inv_rowsum.numel(), n, row_offset, mode, int(use_gsum), part, world, rank, _p(gvec),
                           _p(scale_dev), _p(wr), _p(wc), _p(dg), _p(out_scale_a), _p(out_scale_b), what)


def dz_panel(A_rows, B_all, grow0, scale_dev, stats, wr, wc, dg, Wz):
    CALLS.append("dz_panel")
    rows, d = A_rows.shape
    _tc().emu_dz_panel(_p(A_rows), _p(B_all), rows, B_all.shape[0], d, grow0, _p(scale_dev), _p(stats), _p(wr), _p(wc), _p(dg), _p(Wz),
                       Wz.stride(0), 0)


def gemm_rowdot_scratch_floats(M, Nc):
    return 2 * ((Nc + 255) // 256) * ((M + 127) // 128 * 128)


def gemm_bf16(A, a_mn, B, b_mn, M, Nc, K, *, acc_in=None, acc_out=None, out=None, row_scale=None, dot_mat=None, rowdot_part=None):
    CALLS.append("gemm")
    ref = out if out is not None else acc_out
    ldc = ref.stride(0)
    _tc().emu_gemm(_p(A), A.stride(0), int(a_mn), _p(B), B.stride(0), int(b_mn), M, Nc, K, _p(acc_in), _p(acc_out), _p(out), ldc, _p(row_scale),
                   _p(dot_mat), dot_mat.stride(0) if dot_mat is not None else 0, _p(rowdot_part))


def rowdot_bf16(x, y, out):
    CALLS.append("rowdot")
    _vec().emu_rowdot_bf16(_p(x), x.stride(0), _p(y), y.stride(0), x.shape[0], x.shape[1], _p(out))


def sum_f32(v, out):
    CALLS.append("sum")
    _vec().emu_sum_f32(_p(v), v.numel(), _p(out))


def split_fp32(x, out, side, terms):
    CALLS.append("split")
    _vec().emu_split_fp32(_p(x), _p(out), x.shape[0], x.shape[1], side, terms)


def rowcol_max(A, B_all, scale_dev, rowmax, colmax, scratch=None):
    CALLS.append("rowcol_max")
    n, d = A.shape
    N = B_all.shape[0]
    s = _scratch(n, N)         # keep the scratch alive across the call
    _tc().emu_rowcol_max(_p(A), _p(B_all), n, N, d, _p(scale_dev), _p(rowmax), _p(colmax), _p(s))
    return scratch


def augment(x, ref, scale_dev, out, ref_q=None):
    CALLS.append("augment")
    _vec().emu_augment(_p(x), x.shape[0], x.shape[1], _p(ref), _p(scale_dev), _p(out), _p(ref_q))


def retrieval_ranks(S, M, label_dot, rank_s2m, rank_m2s, scratch=None):
    CALLS.append("retrieval_ranks")
    N, d = S.shape
    s = _scratch(N, N)
    _tc().emu_retrieval_ranks(_p(S), _p(M), N, d, _p(label_dot), _p(rank_s2m), _p(rank_m2s), _p(s))


def siglip_fwd(A, B_all, scale_dev, bias_dev, rowsum, scratch=None):
    CALLS.append("siglip_fwd")
    n, d = A.shape
    N = B_all.shape[0]
    s = _scratch(n, N)
    _tc().emu_siglip_fwd(_p(A), _p(B_all), n, N, d, _p(scale_dev), _p(bias_dev), _p(rowsum), _p(s))
    return scratch


def siglip_fwd_keep(A, B_all, grow0, scale_dev, bias_dev, rowsum, S, sig_rowsum=None):
    CALLS.append("siglip_fwd_keep")
    n, d = A.shape
    N = B_all.shape[0]
    s1, s2 = _scratch(n, N), _scratch(n, N)
    _tc().emu_siglip_fwd_keep(_p(A), _p(B_all), n, N, d, grow0, _p(scale_dev), _p(bias_dev), _p(rowsum), _p(sig_rowsum), _p(s1), _p(s2),
                              _p(S), S.stride(0))


def siglip_finalize(rowsum, diag, scale_dev, bias_dev, loss_out):
    CALLS.append("siglip_finalize")
    _vec().emu_siglip_finalize(_p(rowsum), _p(diag), rowsum.numel(), _p(scale_dev), _p(bias_dev), _p(loss_out))


def siglip_dz_panel(A_rows, B_all, grow0, scale_dev, bias_dev, wr, dg, Wz, sig_rowsum=None):
    CALLS.append("siglip_dz_panel")
    rows, d = A_rows.shape
    s = _scratch(rows, B_all.shape[0])
    _tc().emu_siglip_dz(_p(A_rows), _p(B_all), rows, B_all.shape[0], d, grow0, _p(scale_dev), _p(bias_dev), _p(wr), _p(dg), _p(Wz), Wz.stride(0),
                        _p(sig_rowsum), _p(s))


# ---- epilogue + heads ------------------------------------------------------------------------------
def l2norm_scale_fwd(x, y, inv_norm, scale_dev=None, eps=1e-12):
    CALLS.append("l2norm_fwd")
    _vec().emu_l2norm_fwd(_p(x), _p(y), _p(inv_norm), x.shape[0], x.shape[1], _f32(x), _p(scale_dev), C.c_float(eps))


def l2norm_scale_bwd(x, gy, inv_norm, gx, dscale_partial=None, scale_dev=None, eps=1e-12):
    CALLS.append("l2norm_bwd")
    _vec().emu_l2norm_bwd(_p(x), _p(gy), _p(inv_norm), _p(gx), _p(dscale_partial), x.shape[0], x.shape[1], _f32(x), _p(scale_dev), C.c_float(eps))


def scale_rows(x, y, scale_dev):
    CALLS.append("scale_rows")
    _vec().emu_scale_rows(_p(x), _p(y), x.shape[0], x.shape[1], _f32(x), _p(scale_dev))


def rowdot(x, y, out):
    CALLS.append("rowdot_dense")
    _vec().emu_rowdot(_p(x), _p(y), x.shape[0], x.shape[1], _f32(x), _p(out))


def layernorm_fwd(x, gamma, beta, y, mean, rstd, eps):
    CALLS.append("layernorm_fwd")
    _head().emu_layernorm_fwd(_p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), x.shape[0], x.shape[1], _f32(x), C.c_float(eps))


def layernorm_bwd(x, gy, gamma, mean, rstd, gx=None, dgamma=None, dbeta=None):
    CALLS.append("layernorm_bwd")
    rows, d = x.shape
    tmp_gx = gx if gx is not None else torch.empty_like(x)
    tmp_dg = dgamma if dgamma is not None else torch.empty(d)
    tmp_db = dbeta if dbeta is not None else torch.empty(d)
    scratch = torch.empty(int(_head().emu_ln_scratch_floats(rows, d)))
    _head().emu_layernorm_bwd(_p(x), _p(gy), _p(gamma), _p(mean), _p(rstd), _p(tmp_gx), _p(tmp_dg), _p(tmp_db), _p(scratch), rows, d, _f32(x))


def gelu(x, out, gy=None):
    CALLS.append("gelu")
    _head().emu_gelu(_p(x), _p(gy), _p(out), C.c_size_t(x.numel()), _f32(x))


def abs_mean_fwd(x, true_count, out):
    CALLS.append("abs_mean_fwd")
    part = torch.zeros(4096, dtype=torch.float32)
    _head().emu_abs_mean_fwd(_p(x), C.c_size_t(x.numel()), C.c_size_t(int(true_count)), _f32(x), _p(out), _p(part))


def abs_mean_bwd(x, g, true_count, gx):
    CALLS.append("abs_mean_bwd")
    _head().emu_abs_mean_bwd(_p(x), _p(g), C.c_size_t(x.numel()), C.c_size_t(int(true_count)), _f32(x), _p(gx))


def meanpool_fwd(x, mask, y, inv_count, normalize=True):
    CALLS.append("meanpool_fwd")
    B, L, D = x.shape
    _head().emu_meanpool_fwd(_p(x), _p(mask), _p(y), _p(inv_count), B, L, D, _f32(x), int(normalize))


def meanpool_bwd(gy, mask, inv_count, gx):
    CALLS.append("meanpool_bwd")
    B, L, D = gx.shape
    _head().emu_meanpool_bwd(_p(gy), _p(mask), _p(inv_count), _p(gx), B, L, D, _f32(gx))


def token_dot(x, vec, out, bias=None, mask=None):
    CALLS.append("token_dot")
    B, L, D = x.shape
    _head().emu_token_dot(_p(x), _p(vec), int(vec.dim() == 2), _p(bias), _p(mask), _p(out), B, L, D, _f32(x))


def softmax_rows(s, p):
    CALLS.append("softmax_rows")
    _head().emu_softmax_rows(_p(s), _p(p), s.shape[0], s.shape[1])


def softmax_rows_bwd(p, dp, ds):
    CALLS.append("softmax_rows_bwd")
    _head().emu_softmax_rows_bwd(_p(p), _p(dp), _p(ds), p.shape[0], p.shape[1])


def attnpool_bwd_x(g, p, ds, w, gx):
    CALLS.append("attnpool_bwd_x")
    B, L, D = gx.shape
    _head().emu_attnpool_bwd_x(_p(g), _p(p), _p(ds), _p(w), _p(gx), B, L, D, _f32(gx))


def sum_slots_f32(part, out):
    CALLS.append("sum_slots_f32")
    out.copy_(part.double().sum(0).float())       # the device kernel is exercised in test_kernel_emulation_cpu.py
