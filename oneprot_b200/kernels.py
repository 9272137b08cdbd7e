"""Tensor-level wrappers over the C ABI (include/oneprot_clip.h).

PyTorch is used here only for device memory and the current CUDA stream; every computation is a
call into liboneprot_clip.so.  No fallbacks: a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr

MODE_GLOBAL = 0
MODE_LOCAL = 1


import threading

_TLS = threading.local()


class stream_scope:
    """Caches the raw handle of torch's current CUDA stream for the launches issued inside the
    `with` block (one lookup per forward / backward instead of one per kernel).  Nested scopes
    (e.g. a side stream) restore the outer handle on exit."""

    def __enter__(self):
        self.prev = getattr(_TLS, "stream", None)
        _TLS.stream = "unset"          # looked up lazily by the first launch inside the scope
        return self

    def __exit__(self, *exc):
        _TLS.stream = self.prev
        return False


def current_stream_handle() -> int:
    """Raw cudaStream_t of torch's current stream (dry-run traces: the test's stand-in)."""
    if _DRY is not None:
        return int(_DRY())
    return int(torch.cuda.current_stream().cuda_stream)


def _stream() -> C.c_void_p:
    s = getattr(_TLS, "stream", None)
    if s is None:
        return C.c_void_p(current_stream_handle())
    if s == "unset":
        s = _TLS.stream = C.c_void_p(current_stream_handle())
    return s


# Launch-trace dry run (tests/test_sequencer_cpu.py): a callable returning the stand-in handle of the
# "current stream".  While it is set the library skips every CUDA call (oneprot_trace_begin(1)), so
# the wrappers accept CPU tensors - their addresses only label the trace lines.
_DRY = None


class launch_trace:
    """Records one text line per entry-point call of liboneprot_clip.so; ``dry_stream`` (a callable
    returning the current stand-in stream handle) makes it a dry run without any CUDA work."""

    def __init__(self, dry_stream=None):
        self.dry_stream = dry_stream
        self.lines = []

    def __enter__(self):
        global _DRY
        _lib.load().oneprot_trace_begin(1 if self.dry_stream is not None else 0)
        _DRY = self.dry_stream
        return self

    def __exit__(self, *exc):
        global _DRY
        _DRY = None
        lib = _lib.load()
        n = int(lib.oneprot_trace_end(None, 0))
        buf = C.create_string_buffer(n + 1)
        lib.oneprot_trace_end(buf, n + 1)
        self.lines = [ln for ln in buf.value.decode().split("\n") if ln]
        return False


def trace_note(text: str):
    _lib.load().oneprot_trace_note(text.encode())


def _need_cuda(*ts):
    if _DRY is not None:
        return
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.OneProtKernelError("oneprot_b200 kernels need CUDA tensors (no CPU fallback exists)")


def _need(t, dtype, what):
    if t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f"{what}: expected contiguous {dtype}, got {t.dtype} contiguous={t.is_contiguous()}")


def panel_row_unit(d: int) -> int:
    """Smallest panel height (rows) whose dA GEMM tile count (128 x 256 tiles) is a multiple of the
    SM count, i.e. fills whole waves of the persistent grid."""
    import math
    sms = int(_lib.load().oneprot_num_sms())
    n_col_blocks = (d + 255) // 256
    return 128 * (sms // math.gcd(sms, n_col_blocks))


def launch_count() -> int:
    return int(_lib.load().oneprot_launch_count())


def launch_count_reset():
    _lib.load().oneprot_launch_count_reset()


def _need_stats(stats):
    """[max |a|^2, max |b|^2, exact maximum logit, its valid flag]: the forward / panel kernels read all four floats."""
    _need(stats, torch.float32, "stats")
    if stats.numel() < 4:
        raise ValueError("stats must hold 4 floats: [max|a|^2, max|b|^2, exact max logit, valid flag]")


def rowstats(A, B_all, row_offset: int, diag, stats):
    """diag[i] = <a_i, b_{row_offset+i}>, stats[0:2] = max |a|^2, max |b|^2 (atomic max)."""
    _need_cuda(A, B_all, diag, stats)
    _need(A, torch.bfloat16, "A"); _need(B_all, torch.bfloat16, "B_all")
    _need(diag, torch.float32, "diag"); _need_stats(stats)
    n, d = A.shape
    N = B_all.shape[0]
    check(_lib.load().oneprot_clip_rowstats(ptr(A), ptr(B_all), n, N, d, row_offset, ptr(diag), ptr(stats),
                                            _stream()), "oneprot_clip_rowstats")


def fwd_scratch_bytes(n: int, N: int) -> int:
    return int(_lib.load().oneprot_clip_fwd_scratch_bytes(n, N))


def fwd_sums(A, B_all, scale_dev, stats, rowsum, colsum, scratch=None, ag=None, keep=None):
    """rowsum[i] = sum_j e_ij, colsum[j] = sum_i e_ij for the n x N logit panel (never stored).
    ag: optional dict describing the fused all-gather (see oneprot_ag_t).
    keep: optional bf16 tensor (>= n rows, row pitch >= N, multiple of 8) that receives the exponentials
    e_ij for dz_from_exp (stored-exponentials backward)."""
    _need_cuda(A, B_all, scale_dev, stats, rowsum, colsum)
    _need(A, torch.bfloat16, "A"); _need(B_all, torch.bfloat16, "B_all"); _need_stats(stats)
    n, d = A.shape
    N = B_all.shape[0]
    need = fwd_scratch_bytes(n, N)
    if scratch is None or scratch.numel() * scratch.element_size() < need:
        scratch = torch.empty(need, dtype=torch.uint8, device=A.device)
    desc = None
    if ag is not None:
        desc = _lib.AgDesc(ag["src"], ag["dst_mc"], ag["counters"], ag["flags_mc"], ag["flags"], ag["stats_mc"],
                           ag["stats_all"], ag["stats_out"], ag["epoch"], ag["rank"], ag["world"], ag["chunks"],
                           ag["rows_per_rank"])
    if keep is not None:
        _need_cuda(keep)
        _need(keep, torch.bfloat16, "keep")
        if keep.dim() != 2 or keep.shape[0] < n or keep.stride(1) != 1:
            raise ValueError("keep must be a row-major 2-D tensor with at least n rows")
        check(_lib.load().oneprot_clip_fwd_sums_keep(ptr(A), ptr(B_all), n, N, d, ptr(scale_dev), ptr(stats),
                                                     C.byref(desc) if desc is not None else None, ptr(rowsum), ptr(colsum),
                                                     ptr(scratch), scratch.numel() * scratch.element_size(), ptr(keep),
                                                     keep.stride(0), _stream()),
              "oneprot_clip_fwd_sums_keep")
        return scratch
    check(_lib.load().oneprot_clip_fwd_sums_ag(ptr(A), ptr(B_all), n, N, d, ptr(scale_dev), ptr(stats),
                                               C.byref(desc) if desc is not None else None, ptr(rowsum), ptr(colsum),
                                               ptr(scratch), scratch.numel() * scratch.element_size(), _stream()),
          "oneprot_clip_fwd_sums_ag")
    return scratch


_FIN_SCRATCH = {}   # device index -> zero-initialised scratch of oneprot_clip_loss_finalize


def loss_finalize(rowsum_all, colsum_all, diag_all, n: int, row_offset: int, mode: int, scale_dev, stats,
                  loss_out, inv_rowsum, inv_colsum, flag, row_ref=None, col_ref=None):
    """row_ref / col_ref: per-element references (log2 units) of the two-reference path."""
    _need_cuda(rowsum_all, colsum_all, diag_all, loss_out, inv_rowsum, inv_colsum, flag, row_ref, col_ref)
    N = rowsum_all.numel()
    key = (rowsum_all.device.index, current_stream_handle())
    scratch = _FIN_SCRATCH.get(key)
    if scratch is None:
        scratch = _FIN_SCRATCH[key] = torch.zeros(64, dtype=torch.float64, device=rowsum_all.device)
    check(_lib.load().oneprot_clip_loss_finalize_ex(ptr(rowsum_all), ptr(colsum_all), ptr(diag_all), N, n, row_offset,
                                                    mode, ptr(scale_dev), ptr(stats), ptr(loss_out), ptr(inv_rowsum),
                                                    ptr(inv_colsum), ptr(flag), ptr(scratch), ptr(row_ref), ptr(col_ref),
                                                    _stream()),
          "oneprot_clip_loss_finalize_ex")


def rowcol_max(A, B_all, scale_dev, rowmax, colmax, scratch=None):
    """Per-row maxima (complete) and per-column maxima over these rows of x = c <a_i, b_j> (log2 units)."""
    _need_cuda(A, B_all, scale_dev, rowmax, colmax)
    _need(A, torch.bfloat16, "A"); _need(B_all, torch.bfloat16, "B_all")
    n, d = A.shape
    N = B_all.shape[0]
    need = fwd_scratch_bytes(n, N)
    if scratch is None or scratch.numel() * scratch.element_size() < need:
        scratch = torch.empty(need, dtype=torch.uint8, device=A.device)
    check(_lib.load().oneprot_clip_rowcol_max(ptr(A), ptr(B_all), n, N, d, ptr(scale_dev), ptr(rowmax), ptr(colmax),
                                              ptr(scratch), scratch.numel() * scratch.element_size(), _stream()),
          "oneprot_clip_rowcol_max")
    return scratch


def augment(x, ref, scale_dev, out, ref_q=None):
    """out = [x | e_h | e_m | 0 x 6] with e_h + e_m = -ref/c (two bf16 limbs) or 1, 1; ref_q = reference actually applied."""
    _need_cuda(x, ref, scale_dev, out, ref_q)
    _need(x, torch.bfloat16, "x"); _need(out, torch.bfloat16, "out")
    rows, d = x.shape
    if out.shape != (rows, d + 8):
        raise ValueError("augment: out must be rows x (d + 8)")
    check(_lib.load().oneprot_augment_bf16(ptr(x), rows, d, ptr(ref), ptr(scale_dev), ptr(out), ptr(ref_q), _stream()),
          "oneprot_augment_bf16")


def bwd_weights(inv_rowsum, inv_colsum, n: int, row_offset: int, mode: int, use_gsum: bool, part: int, world: int,
                rank: int, gvec, scale_dev, wr, wc, dg, out_scale_a, out_scale_b, what: int = 0):
    """what: 0 = everything, 1 = panel weights (wr, wc, dg) only, 2 = output scales only."""
    _need_cuda(inv_rowsum, inv_colsum, gvec, scale_dev, wr, wc, dg, out_scale_a, out_scale_b)
    N = inv_rowsum.numel()
    check(_lib.load().oneprot_clip_bwd_weights(ptr(inv_rowsum), ptr(inv_colsum), N, n, row_offset, mode,
                                               int(use_gsum), part, world, rank, ptr(gvec), ptr(scale_dev), ptr(wr),
                                               ptr(wc), ptr(dg), ptr(out_scale_a), ptr(out_scale_b), what, _stream()),
          "oneprot_clip_bwd_weights")


def dz_panel(A_rows, B_all, grow0: int, scale_dev, stats, wr, wc, dg, Wz):
    """Wz[i, j] = e_ij (wr[i] + wc[j]) - [grow0+i == j] dg[i]  (bf16 panel, rows x ldw)."""
    _need_cuda(A_rows, B_all, wr, wc, dg, Wz)
    _need(A_rows, torch.bfloat16, "A_rows"); _need(B_all, torch.bfloat16, "B_all"); _need(Wz, torch.bfloat16, "Wz")
    _need_stats(stats)
    rows, d = A_rows.shape
    N = B_all.shape[0]
    if Wz.shape[0] < rows:
        raise ValueError("Wz panel has fewer rows than A_rows")
    check(_lib.load().oneprot_clip_dz_panel(ptr(A_rows), ptr(B_all), rows, N, d, grow0, ptr(scale_dev), ptr(stats),
                                            ptr(wr), ptr(wc), ptr(dg), ptr(Wz), Wz.stride(0), _stream()),
          "oneprot_clip_dz_panel")


def dz_from_exp(E, rows: int, N: int, grow0: int, wr, wc, dg):
    """In place on the stored exponentials: E[i, j] <- E[i, j] (wr[i] + wc[j]) - [grow0+i == j] dg[i] for
    i < rows, j < N - the dL/dZ panel of dz_panel without a second pass over the logits."""
    _need_cuda(E, wr, wc, dg)
    _need(E, torch.bfloat16, "E")
    if E.dim() != 2 or E.shape[0] < rows or E.stride(1) != 1:
        raise ValueError("E must be a row-major 2-D tensor with at least `rows` rows")
    check(_lib.load().oneprot_clip_dz_from_exp(ptr(E), rows, N, E.stride(0), grow0, ptr(wr), ptr(wc), ptr(dg), _stream()),
          "oneprot_clip_dz_from_exp")


def siglip_fwd(A, B_all, scale_dev, bias_dev, rowsum, scratch=None):
    """rowsum[i] = sum_j log2(1 + 2^x_ij), x = log2(e) (scale <a_i, b_j> + bias)  (SigLIP forward)."""
    _need_cuda(A, B_all, scale_dev, bias_dev, rowsum)
    _need(A, torch.bfloat16, "A"); _need(B_all, torch.bfloat16, "B_all")
    n, d = A.shape
    N = B_all.shape[0]
    need = fwd_scratch_bytes(n, N)
    if scratch is None or scratch.numel() * scratch.element_size() < need:
        scratch = torch.empty(need, dtype=torch.uint8, device=A.device)
    check(_lib.load().oneprot_siglip_fwd(ptr(A), ptr(B_all), n, N, d, ptr(scale_dev), ptr(bias_dev), ptr(rowsum), ptr(scratch),
                                         scratch.numel() * scratch.element_size(), _stream()), "oneprot_siglip_fwd")
    return scratch


def siglip_fwd_keep(A, B_all, grow0: int, scale_dev, bias_dev, rowsum, S, sig_rowsum=None):
    """siglip_fwd that also keeps S[i, j] = sigma(z_ij) - [grow0+i == j] (bf16, >= n rows, row pitch >= N) and, optionally,
    the row sums of sigma."""
    _need_cuda(A, B_all, scale_dev, bias_dev, rowsum, S, sig_rowsum)
    _need(A, torch.bfloat16, "A"); _need(B_all, torch.bfloat16, "B_all"); _need(S, torch.bfloat16, "S")
    n, d = A.shape
    N = B_all.shape[0]
    if S.dim() != 2 or S.shape[0] < n or S.stride(1) != 1:
        raise ValueError("S must be a row-major 2-D tensor with at least n rows")
    lib = _lib.load()
    need = int(lib.oneprot_siglip_fwd_keep_scratch_bytes(n, N))
    scratch = torch.empty(need, dtype=torch.uint8, device=A.device)
    check(lib.oneprot_siglip_fwd_keep(ptr(A), ptr(B_all), n, N, d, grow0, ptr(scale_dev), ptr(bias_dev), ptr(rowsum), ptr(sig_rowsum),
                                      ptr(scratch), need, ptr(S), S.stride(0), _stream()), "oneprot_siglip_fwd_keep")


def siglip_finalize(rowsum, diag, scale_dev, bias_dev, loss_out):
    _need_cuda(rowsum, diag, scale_dev, bias_dev, loss_out)
    check(_lib.load().oneprot_siglip_finalize(ptr(rowsum), ptr(diag), rowsum.numel(), ptr(scale_dev), ptr(bias_dev), ptr(loss_out),
                                              _stream()), "oneprot_siglip_finalize")


def siglip_dz_panel(A_rows, B_all, grow0: int, scale_dev, bias_dev, wr, dg, Wz, sig_rowsum=None):
    """Wz[i, j] = wr[i] sigma(z_ij) - [grow0 + i == j] dg[i]  (bf16 panel, rows x ldw);
    sig_rowsum (optional, fp32[rows]): sum_j sigma(z_ij)."""
    _need_cuda(A_rows, B_all, scale_dev, bias_dev, wr, dg, Wz, sig_rowsum)
    _need(A_rows, torch.bfloat16, "A_rows"); _need(B_all, torch.bfloat16, "B_all"); _need(Wz, torch.bfloat16, "Wz")
    rows, d = A_rows.shape
    N = B_all.shape[0]
    if Wz.shape[0] < rows:
        raise ValueError("Wz panel has fewer rows than A_rows")
    scratch, nbytes = None, 0
    if sig_rowsum is not None:
        nbytes = int(_lib.load().oneprot_siglip_dz_scratch_bytes(rows, N))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=A_rows.device)
    check(_lib.load().oneprot_siglip_dz_panel(ptr(A_rows), ptr(B_all), rows, N, d, grow0, ptr(scale_dev), ptr(bias_dev), ptr(wr),
                                              ptr(dg), ptr(Wz), Wz.stride(0), ptr(sig_rowsum), ptr(scratch), nbytes, _stream()),
          "oneprot_siglip_dz_panel")


def gemm_rowdot_scratch_floats(M: int, Nc: int) -> int:
    return int(_lib.load().oneprot_gemm_rowdot_scratch_bytes(M, Nc)) // 4


def gemm_bf16(A, a_mn: bool, B, b_mn: bool, M: int, Nc: int, K: int, *, acc_in=None, acc_out=None, out=None,
              row_scale=None, dot_mat=None, rowdot_part=None):
    """C[M x Nc] = op(A) op(B); see oneprot_gemm_bf16(_ex) in include/oneprot_clip.h."""
    _need_cuda(A, B, acc_in, acc_out, out, row_scale, dot_mat, rowdot_part)
    if A.dtype != torch.bfloat16 or B.dtype != torch.bfloat16:
        raise ValueError("gemm_bf16 needs bf16 operands")
    if A.stride(1) != 1 or B.stride(1) != 1:
        raise ValueError("gemm_bf16 operands must be row-major (unit inner stride)")
    ref = out if out is not None else acc_out
    ldc = ref.stride(0)
    for t in (acc_in, acc_out, out):
        if t is not None and (t.stride(0) != ldc or t.stride(1) != 1):
            raise ValueError("gemm_bf16 outputs must share one row-major leading dimension")
    ld_dot = 0
    if dot_mat is not None:
        if dot_mat.dtype != torch.bfloat16 or dot_mat.stride(1) != 1:
            raise ValueError("dot_mat must be row-major bf16")
        ld_dot = dot_mat.stride(0)
    check(_lib.load().oneprot_gemm_bf16_ex(ptr(A), A.stride(0), int(a_mn), ptr(B), B.stride(0), int(b_mn), M, Nc, K,
                                           ptr(acc_in), ptr(acc_out), ptr(out), ldc, ptr(row_scale), ptr(dot_mat),
                                           ld_dot, ptr(rowdot_part), _stream()), "oneprot_gemm_bf16_ex")


def rowdot_bf16(x, y, out):
    _need_cuda(x, y, out)
    rows, d = x.shape
    check(_lib.load().oneprot_rowdot_bf16(ptr(x), x.stride(0), ptr(y), y.stride(0), rows, d, ptr(out), _stream()),
          "oneprot_rowdot_bf16")


def sum_f32(v, out):
    _need_cuda(v, out)
    check(_lib.load().oneprot_sum_f32(ptr(v), v.numel(), ptr(out), _stream()), "oneprot_sum_f32")


def l2norm_scale_fwd(x, y, inv_norm, scale_dev=None, eps: float = 1e-12):
    _need_cuda(x, y, inv_norm, scale_dev)
    rows, d = x.shape
    check(_lib.load().oneprot_l2norm_scale_fwd(ptr(x), ptr(y), ptr(inv_norm), rows, d,
                                               int(x.dtype == torch.float32), ptr(scale_dev), eps, _stream()),
          "oneprot_l2norm_scale_fwd")


def l2norm_scale_bwd(x, gy, inv_norm, gx, dscale_partial=None, scale_dev=None, eps: float = 1e-12):
    _need_cuda(x, gy, inv_norm, gx, dscale_partial, scale_dev)
    rows, d = x.shape
    check(_lib.load().oneprot_l2norm_scale_bwd(ptr(x), ptr(gy), ptr(inv_norm), ptr(gx), ptr(dscale_partial), rows, d,
                                               int(x.dtype == torch.float32), ptr(scale_dev), eps, _stream()),
          "oneprot_l2norm_scale_bwd")


def split_fp32(x, out, side: int, terms: int):
    _need_cuda(x, out)
    rows, d = x.shape
    check(_lib.load().oneprot_split_fp32(ptr(x), ptr(out), rows, d, side, terms, _stream()), "oneprot_split_fp32")


def scale_rows(x, y, scale_dev):
    _need_cuda(x, y, scale_dev)
    rows, d = x.shape
    check(_lib.load().oneprot_scale_rows(ptr(x), ptr(y), rows, d, int(x.dtype == torch.float32), ptr(scale_dev),
                                         _stream()), "oneprot_scale_rows")


def rowdot(x, y, out):
    _need_cuda(x, y, out)
    rows, d = x.shape
    check(_lib.load().oneprot_rowdot(ptr(x), ptr(y), rows, d, int(x.dtype == torch.float32), ptr(out), _stream()),
          "oneprot_rowdot")


# ---- projection-head row kernels (pooling, LayerNorm, GELU) ---------------------------------
def _is32(t):
    if t.dtype not in (torch.bfloat16, torch.float32) or not t.is_contiguous():
        raise ValueError(f"expected a contiguous bf16 / fp32 tensor, got {t.dtype} contiguous={t.is_contiguous()}")
    return int(t.dtype == torch.float32)


def layernorm_fwd(x, gamma, beta, y, mean, rstd, eps: float):
    _need_cuda(x, gamma, beta, y, mean, rstd)
    rows, d = x.shape
    if gamma.dtype != x.dtype or beta.dtype != x.dtype:
        raise ValueError("layernorm: gamma / beta must have the dtype of x")
    check(_lib.load().oneprot_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), rows, d, _is32(x),
                                            eps, _stream()), "oneprot_layernorm_fwd")


def layernorm_bwd(x, gy, gamma, mean, rstd, gx=None, dgamma=None, dbeta=None):
    _need_cuda(x, gy, gamma, mean, rstd, gx, dgamma, dbeta)
    rows, d = x.shape
    scratch, nbytes = None, 0
    if dgamma is not None:
        nbytes = int(_lib.load().oneprot_layernorm_bwd_scratch_bytes(rows, d))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    if gy.dtype != x.dtype or not gy.is_contiguous():
        raise ValueError("layernorm_bwd: gy must be contiguous with the dtype of x")
    check(_lib.load().oneprot_layernorm_bwd(ptr(x), ptr(gy), ptr(gamma), ptr(mean), ptr(rstd), ptr(gx), ptr(dgamma), ptr(dbeta),
                                            ptr(scratch), nbytes, rows, d, _is32(x), _stream()), "oneprot_layernorm_bwd")


def gelu(x, out, gy=None):
    """out = gelu(x), or gy * gelu'(x) when gy is given (flat, element count a multiple of 8)."""
    _need_cuda(x, out, gy)
    check(_lib.load().oneprot_gelu(ptr(x), ptr(gy), ptr(out), x.numel(), _is32(x), _stream()), "oneprot_gelu")


def meanpool_fwd(x, mask, y, inv_count, normalize: bool = True):
    """y[b] = sum_l mask[b,l] x[b,l,:] (/ sum_l mask[b,l] when normalize)."""
    _need_cuda(x, mask, y, inv_count)
    B, L, D = x.shape
    check(_lib.load().oneprot_meanpool_fwd(ptr(x), ptr(mask), ptr(y), ptr(inv_count), B, L, D, _is32(x), int(normalize),
                                           _stream()), "oneprot_meanpool_fwd")


def token_dot(x, vec, out, bias=None, mask=None):
    """out[b,l] = <vec, x[b,l,:]> + bias; vec: (D,) shared or (B, D) per batch row; -inf where mask == 0."""
    _need_cuda(x, vec, out, bias, mask)
    B, L, D = x.shape
    if vec.dtype != x.dtype or not vec.is_contiguous():
        raise ValueError("token_dot: vec must be contiguous with the dtype of x")
    check(_lib.load().oneprot_token_dot(ptr(x), ptr(vec), int(vec.dim() == 2), ptr(bias), ptr(mask), ptr(out), B, L, D, _is32(x),
                                        _stream()), "oneprot_token_dot")


def softmax_rows(s, p):
    _need_cuda(s, p)
    B, L = s.shape
    check(_lib.load().oneprot_softmax_rows(ptr(s), ptr(p), B, L, _stream()), "oneprot_softmax_rows")


def softmax_rows_bwd(p, dp, ds):
    _need_cuda(p, dp, ds)
    B, L = p.shape
    check(_lib.load().oneprot_softmax_rows_bwd(ptr(p), ptr(dp), ptr(ds), B, L, _stream()), "oneprot_softmax_rows_bwd")


def attnpool_bwd_x(g, p, ds, w, gx):
    _need_cuda(g, p, ds, w, gx)
    B, L, D = gx.shape
    check(_lib.load().oneprot_attnpool_bwd_x(ptr(g), ptr(p), ptr(ds), ptr(w), ptr(gx), B, L, D, _is32(gx), _stream()),
          "oneprot_attnpool_bwd_x")


def abs_mean_fwd(x, true_count: int, out):
    """out[0] = sum |x| / true_count (x flat, element count a multiple of 8, zero padded beyond true_count)."""
    _need_cuda(x, out)
    lib = _lib.load()
    nbytes = int(lib.oneprot_abs_mean_scratch_bytes(x.numel()))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    check(lib.oneprot_abs_mean_fwd(ptr(x), x.numel(), int(true_count), _is32(x), ptr(out), ptr(scratch), nbytes, _stream()),
          "oneprot_abs_mean_fwd")


def abs_mean_bwd(x, g, true_count: int, gx):
    """gx = g[0] sign(x) / true_count."""
    _need_cuda(x, g, gx)
    check(_lib.load().oneprot_abs_mean_bwd(ptr(x), ptr(g), x.numel(), int(true_count), _is32(x), ptr(gx), _stream()),
          "oneprot_abs_mean_bwd")


def sum_slots_f32(part, out):
    """out[k] = sum_s part[s, k] (part: slots x count fp32, fixed order)."""
    _need_cuda(part, out)
    slots, count = part.shape
    check(_lib.load().oneprot_sum_slots_f32(ptr(part), slots, part.stride(0), count, ptr(out), _stream()),
          "oneprot_sum_slots_f32")


def meanpool_bwd(gy, mask, inv_count, gx):
    _need_cuda(gy, mask, inv_count, gx)
    B, L, D = gx.shape
    check(_lib.load().oneprot_meanpool_bwd(ptr(gy), ptr(mask), ptr(inv_count), ptr(gx), B, L, D, _is32(gx), _stream()),
          "oneprot_meanpool_bwd")


# ---- NVLS (multimem) exchanges; *_mc arguments are raw multicast addresses (int) -------------
def mc_store(src, dst_mc_addr: int, nbytes: int):
    _need_cuda(src)
    check(_lib.load().oneprot_mc_store(ptr(src), C.c_void_p(dst_mc_addr), nbytes, _stream()), "oneprot_mc_store")


def mc_allreduce_f32(src_mc_addr: int, dst, count: int, op: int = 0):
    _need_cuda(dst)
    check(_lib.load().oneprot_mc_allreduce_f32(C.c_void_p(src_mc_addr), ptr(dst), count, op, _stream()),
          "oneprot_mc_allreduce_f32")


def mc_reduce_bf16(src_mc_addr: int, dst, nbytes: int):
    _need_cuda(dst)
    check(_lib.load().oneprot_mc_reduce_bf16(C.c_void_p(src_mc_addr), ptr(dst), nbytes, _stream()),
          "oneprot_mc_reduce_bf16")


def retrieval_ranks(S, M, label_dot, rank_s2m, rank_m2s, scratch=None):
    """rank_s2m[i] = #{j != i: <s_i, m_j> > label_dot[i]}, rank_m2s[j] = #{i != j: <s_i, m_j> > label_dot[j]}."""
    _need_cuda(S, M, label_dot, rank_s2m, rank_m2s)
    _need(S, torch.bfloat16, "S"); _need(M, torch.bfloat16, "M")
    N, d = S.shape
    need = fwd_scratch_bytes(N, N)
    if scratch is None or scratch.numel() * scratch.element_size() < need:
        scratch = torch.empty(need, dtype=torch.uint8, device=S.device)
    check(_lib.load().oneprot_retrieval_ranks(ptr(S), ptr(M), N, d, ptr(label_dot), ptr(rank_s2m), ptr(rank_m2s),
                                              ptr(scratch), scratch.numel() * scratch.element_size(), _stream()),
          "oneprot_retrieval_ranks")
