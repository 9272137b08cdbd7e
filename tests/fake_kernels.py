"""CPU emulation of oneprot_b200.kernels for HOST-LOGIC tests only (gloo, world_size 2).

Test infrastructure: lives under tests/, is injected by monkeypatching
``oneprot_b200.clip_loss._KERNELS`` and is never importable from the product package.  Each
function restates the contract of the C entry point of the same name (include/oneprot_clip.h) in
plain torch float64 so that the sharding / collective / gradient-convention logic of
``ClipLoss`` can be exercised without a GPU.
"""
import math

import torch

MODE_GLOBAL = 0
MODE_LOCAL = 1
LOG2E = 1.4426950408889634
CALLS = []


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


def _G(scale_dev, stats):
    c = float(scale_dev[0]) * LOG2E
    U = abs(c) * math.sqrt(float(stats[0]) * float(stats[1]))
    if stats.numel() > 3 and float(stats[3]) != 0.0:        # exact-max / two-reference override
        return c, max(0.0, float(stats[2]) - 100.0)
    return c, max(0.0, U - 100.0)


def rowstats(A, B_all, row_offset, diag, stats):
    CALLS.append("rowstats")
    n = A.shape[0]
    a, b = A.double(), B_all.double()
    diag.copy_((a * b[row_offset:row_offset + n]).sum(-1).float())
    stats[0] = max(float(stats[0]), float((a * a).sum(-1).max()))
    stats[1] = max(float(stats[1]), float((b * b).sum(-1).max()))


def fwd_scratch_bytes(n, N):
    return 16


def fwd_sums(A, B_all, scale_dev, stats, rowsum, colsum, scratch=None, ag=None, keep=None):
    CALLS.append("fwd_sums" if keep is None else "fwd_sums_keep")
    c, G = _G(scale_dev, stats)
    E = torch.exp2(c * (A.double() @ B_all.double().T) - G)
    rowsum.copy_(E.sum(1).float())
    colsum.copy_(E.sum(0).float())
    if keep is not None:
        keep[:A.shape[0], :B_all.shape[0]] = E.float().to(torch.bfloat16)
    return scratch


def dz_from_exp(E, rows, N, grow0, wr, wc, dg):
    CALLS.append("dz_from_exp")
    W = E[:rows, :N].double() * (wr[:rows].double()[:, None] + wc.double()[None, :])
    idx = torch.arange(rows)
    W[idx, grow0 + idx] -= dg[:rows].double()
    E[:rows, :N] = W.float().to(torch.bfloat16)


def loss_finalize(rowsum_all, colsum_all, diag_all, n, row_offset, mode, scale_dev, stats, loss_out, inv_rowsum,
                  inv_colsum, flag, row_ref=None, col_ref=None):
    CALLS.append("loss_finalize")
    c, G = _G(scale_dev, stats)
    s = float(scale_dev[0])
    N = rowsum_all.numel()
    lo, hi = (row_offset, row_offset + n) if mode == MODE_LOCAL else (0, N)
    gr = row_ref.double() if row_ref is not None else G
    gc = col_ref.double() if col_ref is not None else G
    rl = (gr + torch.log2(rowsum_all.double())) / LOG2E
    cl = (gc + torch.log2(colsum_all.double())) / LOG2E
    zd = s * diag_all.double()
    loss_out[0] = float(((rl - zd)[lo:hi].sum() + (cl - zd)[lo:hi].sum()) / (2 * (hi - lo)))
    inv_rowsum.copy_((1.0 / rowsum_all.double()).float())
    inv_colsum.copy_((1.0 / colsum_all.double()).float())
    bad = (~torch.isfinite(rowsum_all)).any() or (~torch.isfinite(colsum_all)).any() \
        or (rowsum_all < 1e-27).any() or (colsum_all < 1e-27).any()
    if bool(bad):
        flag[0] = int(flag[0]) | 1
        loss_out[0] = float("nan")          # outside the validated window the value is returned as NaN


def rowcol_max(A, B_all, scale_dev, rowmax, colmax, scratch=None):
    CALLS.append("rowcol_max")
    c = float(scale_dev[0]) * LOG2E
    X = c * (A.double() @ B_all.double().T)
    rowmax.copy_(X.max(1).values.float())
    colmax.copy_(X.max(0).values.float())
    return scratch


def augment(x, ref, scale_dev, out, ref_q=None):
    CALLS.append("augment")
    c = float(scale_dev[0]) * LOG2E
    d = x.shape[1]
    out.zero_()
    out[:, :d] = x
    if ref is None:
        out[:, d] = 1.0
        out[:, d + 1] = 1.0
    else:
        t = (-ref.double() / c).float()
        eh = t.to(torch.bfloat16)
        em = (t - eh.float()).to(torch.bfloat16)
        out[:, d] = eh
        out[:, d + 1] = em
        if ref_q is not None:
            ref_q.copy_((-c * (eh.float() + em.float())).float())


def bwd_weights(inv_rowsum, inv_colsum, n, row_offset, mode, use_gsum, part, world, rank, gvec, scale_dev, wr, wc, dg,
                out_scale_a, out_scale_b, what=0):
    CALLS.append("bwd_weights")
    if what != 0:   # emulate the selective writes by computing everything into temporaries
        tw = [torch.empty_like(x) for x in (wr, wc, dg, out_scale_a, out_scale_b)]
        bwd_weights(inv_rowsum, inv_colsum, n, row_offset, mode, use_gsum, part, world, rank, gvec, scale_dev, *tw)
        if what == 1:
            wr.copy_(tw[0]); wc.copy_(tw[1]); dg.copy_(tw[2])
        else:
            out_scale_a.copy_(tw[3]); out_scale_b.copy_(tw[4])
        return
    N = inv_rowsum.numel()
    s = float(scale_dev[0])
    g = gvec.double()
    gsum, g_own = float(g.sum()), float(g[rank])
    fr = 0.0 if part == 2 else 1.0
    fc = 0.0 if part == 1 else 1.0
    npr = N // world
    owner = torch.arange(N) // npr
    if mode == MODE_GLOBAL:
        coef = s / (2.0 * N)
        wr.copy_((fr * coef * inv_rowsum[row_offset:row_offset + n].double()).float())
        dg.fill_((fr + fc) * coef)
        out_scale_a.fill_(gsum if use_gsum else g_own)
        wc.copy_((fc * coef * inv_colsum.double()).float())
        out_scale_b.copy_((torch.full((N,), gsum, dtype=torch.float64) if use_gsum else g[owner]).float())
    else:
        coef = s * g_own / (2.0 * n)
        wr.copy_((fr * coef * inv_rowsum[row_offset:row_offset + n].double()).float())
        dg.fill_((fr + fc) * coef)
        out_scale_a.fill_(1.0)
        wc.copy_((fc * s * g[owner] / (2.0 * n) * inv_colsum.double()).float())
        out_scale_b.fill_(1.0)


def dz_panel(A_rows, B_all, grow0, scale_dev, stats, wr, wc, dg, Wz):
    CALLS.append("dz_panel")
    c, G = _G(scale_dev, stats)
    rows, N = A_rows.shape[0], B_all.shape[0]
    E = torch.exp2(c * (A_rows.double() @ B_all.double().T) - G)
    Wd = E * (wr.double()[:, None] + wc.double()[None, :])
    idx = torch.arange(rows)
    Wd[idx, grow0 + idx] -= dg.double()
    Wz[:rows, :N] = Wd.to(torch.bfloat16)


def gemm_rowdot_scratch_floats(M, Nc):
    return 2 * ((Nc + 255) // 256) * ((M + 127) // 128 * 128)


def gemm_bf16(A, a_mn, B, b_mn, M, Nc, K, *, acc_in=None, acc_out=None, out=None, row_scale=None, dot_mat=None,
              rowdot_part=None):
    CALLS.append("gemm")
    opA = (A.double().T if a_mn else A.double())[:M, :K]
    opB = (B.double() if b_mn else B.double().T)[:K, :Nc]
    val = opA @ opB
    if acc_in is not None:
        val = val + acc_in.double()
    if rowdot_part is not None:
        ldd = (M + 127) // 128 * 128
        rp = rowdot_part.view(-1, ldd)
        rp.zero_()
        rp[0, :M] = (val * dot_mat.double()[:M, :Nc]).sum(-1).float()
    if row_scale is not None:
        val = val * row_scale.double()[:M, None]
    if acc_out is not None:
        acc_out.copy_(val.float())
    if out is not None:
        out.copy_(val.to(out.dtype))


def rowdot_bf16(x, y, out):
    CALLS.append("rowdot")
    out.copy_((x.double() * y.double()).sum(-1).float())


def sum_f32(v, out):
    CALLS.append("sum")
    out[0] = float(v.double().sum())


def split_fp32(x, out, side, terms):
    CALLS.append("split")
    h = x.to(torch.bfloat16)
    r1 = x - h.float()
    m = r1.to(torch.bfloat16)
    r2 = r1 - m.float()
    l = r2.to(torch.bfloat16)
    L = [h, h, m, h, m, l]
    R = [h, m, h, l, m, h]
    d = x.shape[1]
    for t in range(terms):
        out[:, t * d:(t + 1) * d] = (R if side else L)[t]


def retrieval_ranks(S, M, label_dot, rank_s2m, rank_m2s, scratch=None):
    CALLS.append("retrieval_ranks")
    Z = S.double() @ M.double().T
    N = Z.shape[0]
    off = ~torch.eye(N, dtype=torch.bool)
    d = label_dot.double()
    rank_s2m.copy_(((Z > d[:, None]) & off).sum(1).float())
    rank_m2s.copy_(((Z > d[None, :]) & off).sum(0).float())


# ---- projection-head row kernels + the normalise / scale epilogue (contracts of include/oneprot_clip.h) ----
def layernorm_fwd(x, gamma, beta, y, mean, rstd, eps):
    CALLS.append("layernorm_fwd")
    xd = x.double()
    mu = xd.mean(-1, keepdim=True)
    var = ((xd - mu) ** 2).mean(-1, keepdim=True)
    rs = 1.0 / torch.sqrt(var + eps)
    y.copy_((((xd - mu) * rs) * gamma.double() + beta.double()).to(y.dtype))
    mean.copy_(mu.squeeze(-1).float())
    rstd.copy_(rs.squeeze(-1).float())


def layernorm_bwd(x, gy, gamma, mean, rstd, gx=None, dgamma=None, dbeta=None):
    CALLS.append("layernorm_bwd")
    xhat = (x.double() - mean.double()[:, None]) * rstd.double()[:, None]
    g = gy.double() * gamma.double()
    if gx is not None:
        gx.copy_((rstd.double()[:, None] * (g - g.mean(-1, keepdim=True) - xhat * (g * xhat).mean(-1, keepdim=True))).to(gx.dtype))
    if dgamma is not None:
        dgamma.copy_((gy.double() * xhat).sum(0).float())
        dbeta.copy_(gy.double().sum(0).float())


def gelu(x, out, gy=None):
    CALLS.append("gelu")
    xd = x.double()
    cdf = 0.5 * (1.0 + torch.erf(xd / math.sqrt(2.0)))
    if gy is None:
        out.copy_((xd * cdf).to(out.dtype))
    else:
        pdf = torch.exp(-0.5 * xd * xd) / math.sqrt(2.0 * math.pi)
        out.copy_((gy.double() * (cdf + xd * pdf)).to(out.dtype))


def meanpool_fwd(x, mask, y, inv_count, normalize=True):
    CALLS.append("meanpool_fwd")
    B, L, D = x.shape
    m = torch.ones(B, L, dtype=torch.float64) if mask is None else mask.double()
    cnt = m.sum(1) if normalize else torch.ones(B, dtype=torch.float64)
    y.copy_(((x.double() * m[:, :, None]).sum(1) / cnt[:, None]).to(y.dtype))
    if inv_count is not None:
        inv_count.copy_((1.0 / cnt).float())


def token_dot(x, vec, out, bias=None, mask=None):
    CALLS.append("token_dot")
    v = vec.double()
    s = (x.double() * (v[:, None, :] if v.dim() == 2 else v)).sum(-1)
    if bias is not None:
        s = s + float(bias[0])
    if mask is not None:
        s = torch.where(mask != 0, s, torch.full_like(s, float("-inf")))
    out.copy_(s.float())


def softmax_rows(s, p):
    CALLS.append("softmax_rows")
    p.copy_(torch.softmax(s.double(), dim=1).float())


def softmax_rows_bwd(p, dp, ds):
    CALLS.append("softmax_rows_bwd")
    pd, dd = p.double(), torch.where(p != 0, dp.double(), torch.zeros_like(dp.double()))
    ds.copy_((pd * (dd - (pd * dd).sum(1, keepdim=True))).float())


def attnpool_bwd_x(g, p, ds, w, gx):
    CALLS.append("attnpool_bwd_x")
    gx.copy_((p.double()[:, :, None] * g.double()[:, None, :] + ds.double()[:, :, None] * w.double()[None, None, :]).to(gx.dtype))


def sum_slots_f32(part, out):
    CALLS.append("sum_slots_f32")
    out.copy_(part.double().sum(0).float())


def meanpool_bwd(gy, mask, inv_count, gx):
    CALLS.append("meanpool_bwd")
    B, L, D = gx.shape
    m = torch.ones(B, L, dtype=torch.float64) if mask is None else mask.double()
    gx.copy_((gy.double()[:, None, :] * (m * inv_count.double()[:, None])[:, :, None]).to(gx.dtype))


def l2norm_scale_fwd(x, y, inv_norm, scale_dev=None, eps=1e-12):
    CALLS.append("l2norm_fwd")
    xd = x.double()
    inv = 1.0 / torch.clamp(xd.norm(dim=-1), min=eps)
    s = 1.0 if scale_dev is None else float(scale_dev[0])
    y.copy_((xd * (inv * s)[:, None]).to(y.dtype))
    inv_norm.copy_(inv.float())


def l2norm_scale_bwd(x, gy, inv_norm, gx, dscale_partial=None, scale_dev=None, eps=1e-12):
    CALLS.append("l2norm_bwd")
    inv = inv_norm.double()
    yhat = x.double() * inv[:, None]
    dot = (yhat * gy.double()).sum(-1)
    if dscale_partial is not None:
        dscale_partial.copy_(dot.float())
    proj = torch.where(inv >= 1.0 / eps, torch.zeros_like(dot), dot)
    s = 1.0 if scale_dev is None else float(scale_dev[0])
    gx.copy_((s * inv[:, None] * (gy.double() - yhat * proj[:, None])).to(gx.dtype))


def scale_rows(x, y, scale_dev):
    CALLS.append("scale_rows")
    y.copy_((x.double() * float(scale_dev[0])).to(y.dtype))


def rowdot(x, y, out):
    CALLS.append("rowdot_dense")
    out.copy_((x.double() * y.double()).sum(-1).float())


# ---- SigLIP (contracts of oneprot_siglip_* in include/oneprot_clip.h) ----
def siglip_fwd(A, B_all, scale_dev, bias_dev, rowsum, scratch=None):
    CALLS.append("siglip_fwd")
    b = 0.0 if bias_dev is None else float(bias_dev[0])
    x = LOG2E * (float(scale_dev[0]) * (A.double() @ B_all.double().T) + b)
    rowsum.copy_((torch.clamp(x, min=0) + torch.log2(1 + torch.exp2(-x.abs()))).sum(1).float())
    return scratch


def siglip_fwd_keep(A, B_all, grow0, scale_dev, bias_dev, rowsum, S, sig_rowsum=None):
    CALLS.append("siglip_fwd_keep")
    b = 0.0 if bias_dev is None else float(bias_dev[0])
    x = LOG2E * (float(scale_dev[0]) * (A.double() @ B_all.double().T) + b)
    rowsum.copy_((torch.clamp(x, min=0) + torch.log2(1 + torch.exp2(-x.abs()))).sum(1).float())
    sg = 1.0 / (1.0 + torch.exp2(-x))
    if sig_rowsum is not None:
        sig_rowsum.copy_(sg.sum(1).float())
    n, N = x.shape
    idx = torch.arange(n)
    sg[idx, grow0 + idx] -= 1.0
    S[:n, :N] = sg.float().to(torch.bfloat16)


def siglip_finalize(rowsum, diag, scale_dev, bias_dev, loss_out):
    CALLS.append("siglip_finalize")
    b = 0.0 if bias_dev is None else float(bias_dev[0])
    n = rowsum.numel()
    loss_out[0] = float((math.log(2.0) * rowsum.double().sum() - (float(scale_dev[0]) * diag.double() + b).sum()) / n)


def siglip_dz_panel(A_rows, B_all, grow0, scale_dev, bias_dev, wr, dg, Wz, sig_rowsum=None):
    CALLS.append("siglip_dz_panel")
    b = 0.0 if bias_dev is None else float(bias_dev[0])
    rows, N = A_rows.shape[0], B_all.shape[0]
    z = float(scale_dev[0]) * (A_rows.double() @ B_all.double().T) + b
    if sig_rowsum is not None:
        sig_rowsum.copy_(torch.sigmoid(z).sum(1).float())
    Wd = torch.sigmoid(z) * wr.double()[:, None]
    idx = torch.arange(rows)
    Wd[idx, grow0 + idx] -= dg.double()
    Wz[:rows, :N] = Wd.to(torch.bfloat16)


def abs_mean_fwd(x, true_count, out):
    CALLS.append("abs_mean_fwd")
    out[0] = float(x.double().abs().sum() / true_count)


def abs_mean_bwd(x, g, true_count, gx):
    CALLS.append("abs_mean_bwd")
    gx.copy_((torch.sign(x.double()) * float(g[0]) / true_count).to(gx.dtype))
