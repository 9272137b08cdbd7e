"""The bar SURVEY.md section 8(d) calls "the real bar": the REFERENCE'S OWN ClipLoss (the unmodified class from
oracle/_ref when it is present, else the oracle's port of loss.py:85-114 - same op sequence) run eagerly by PyTorch
on the same B200: cuBLAS GEMMs + ATen log-softmax / NLL kernels with the N x N logit matrices materialised, autograd
backward.  TEST INFRASTRUCTURE (baseline leg of bench.py and tests/perf_eager_bar.py); never on the product path."""
from __future__ import annotations


def reference_loss_fn():
    """-> (callable(A, B, scale) -> loss, kind)"""
    from oracle.make_ref import import_reference
    ref = import_reference()
    if ref is not None:
        mod = ref[0].ClipLoss(local_loss=False, gather_with_grad=True, cache_labels=True, rank=0, world_size=1)
        return (lambda A, B, s: mod(A, B, s)), "reference (oracle/_ref, unmodified loss.py)"
    from oracle.clip_oracle import clip_loss_port
    return clip_loss_port, "port (oracle.clip_oracle.clip_loss_port)"


def time_eager(a, b, dtype, tf32, dev, flush, reps=5, warm=2):
    """fwd+bwd of the reference op sequence on `dev`; a, b CPU tensors.  -> dict(ms, peak_gib, loss, kind) (ms None on OOM)."""
    import torch
    fn, kind = reference_loss_fn()
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32          # the reference trains with allow_tf32 = True (src/train.py:98)
    A = a.to(dev, dtype).requires_grad_(True)
    B = b.to(dev, dtype).requires_grad_(True)
    out = dict(ms=None, peak_gib=None, loss=None, kind=kind)
    try:
        torch.cuda.reset_peak_memory_stats(dev)
        for _ in range(warm):
            A.grad = None; B.grad = None
            fn(A, B, 1.0).backward()
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(reps):
            flush.zero_()
            A.grad = None; B.grad = None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); loss = fn(A, B, 1.0); loss.backward(); e1.record()
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        out.update(ms=ts[len(ts) // 2], peak_gib=torch.cuda.max_memory_allocated(dev) / 2 ** 30, loss=float(loss.detach().float()))
    except torch.cuda.OutOfMemoryError:
        pass
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
        del A, B
        torch.cuda.empty_cache()
    return out
