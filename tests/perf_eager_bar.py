#!/usr/bin/env python
"""The bar SURVEY.md section 8(d) names besides the CPU baseline: the REFERENCE'S OWN op sequence
(oracle port of loss.py:85-114: scale-then-matmul twice, cross_entropy twice, autograd backward) run
eagerly by PyTorch on the same B200 - cuBLAS GEMMs + ATen softmax / NLL kernels with the N x N logit
matrices materialised - next to this library's fused path, same inputs, CUDA events, L2 flushed.

    python tests/perf_eager_bar.py [--sizes 8192,16384,32768] [--gpus-json gpurun_out/eager_bar.json]
    torchrun --nproc-per-node W tests/perf_eager_bar.py --world            # reference multi-rank port over NCCL

Lives under tests/ because it executes oracle/ (test infrastructure); nothing in the product or
in bench.py's GPU arm calls it."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import clip_oracle as oc  # noqa: E402,F401
from oracle.eager_bar import reference_loss_fn  # noqa: E402
from tools.synthetic import synthetic_pair  # noqa: E402


def timed(fn, reps, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="8192,16384,32768")
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "eager_bar.json"))
    ap.add_argument("--world", action="store_true", help="multi-rank: launch under torchrun, one rank per GPU (NCCL)")
    args = ap.parse_args()
    if args.world:
        return main_world(args)
    from oneprot_b200 import ClipLoss
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for N in [int(x) for x in args.sizes.split(",")]:
        a, b = synthetic_pair(N, args.dim, seed=1234, dtype="bf16")
        for dtype, tf32 in ((torch.bfloat16, False), (torch.float32, True), (torch.float32, False)):
            torch.backends.cuda.matmul.allow_tf32 = tf32      # the reference trains with allow_tf32 = True (src/train.py:98)
            A = a.to(dev, dtype).requires_grad_(True)
            B = b.to(dev, dtype).requires_grad_(True)

            ref_fn, ref_kind = reference_loss_fn()

            def eager():
                A.grad = None; B.grad = None
                ref_fn(A, B, 1.0).backward()

            ours = ClipLoss(loss_dtype=torch.float32)

            def fused():
                A.grad = None; B.grad = None
                ours(A, B).backward()

            try:
                t_eager = timed(eager, args.reps, flush)
                peak = torch.cuda.max_memory_allocated() / 2 ** 30
            except torch.cuda.OutOfMemoryError:
                t_eager, peak = None, None
            torch.cuda.reset_peak_memory_stats()
            t_ours = timed(fused, args.reps, flush)
            peak_ours = torch.cuda.max_memory_allocated() / 2 ** 30
            torch.cuda.reset_peak_memory_stats()
            l_e = float(ref_fn(A.detach(), B.detach(), 1.0)) if t_eager else None
            l_o = float(ours(A.detach(), B.detach()))
            row = dict(N=N, d=args.dim, eager_impl=ref_kind, dtype=str(dtype).split(".")[-1], tf32=tf32, eager_ms=t_eager, fused_ms=t_ours,
                       speedup=(t_eager / t_ours) if t_eager else None, eager_peak_gib=peak, fused_peak_gib=peak_ours,
                       eager_loss=l_e, fused_loss=l_o,
                       eager_samples_per_s=(N / (t_eager * 1e-3)) if t_eager else None, fused_samples_per_s=N / (t_ours * 1e-3))
            print(json.dumps(row), flush=True)
            rows.append(row)
            del A, B
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)


def main_world(args):
    """The reference's multi-rank ClipLoss(local_loss=False, gather_with_grad=True) - BASELINE cfg 2 / the metric's
    mode - eagerly over NCCL: every rank gathers both operands and materialises the full N x N logits (loss.py:95),
    next to this library's sharded path on the same ranks and inputs.  Global N = --sizes, n = N / W rows per rank."""
    import torch.distributed as dist
    from oracle.make_ref import import_reference
    from tools.synthetic import synthetic_global_rows
    from oneprot_b200 import ClipLoss
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", lr)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ref = import_reference()
    rows = []
    for N in [int(x) for x in args.sizes.split(",")]:
        n = N // world
        a, b = synthetic_global_rows(rank * n, n, args.dim, seed=1234, dtype="bf16")
        for dtype, tf32 in ((torch.bfloat16, False), (torch.float32, True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            A = a.to(dev, dtype).requires_grad_(True)
            B = b.to(dev, dtype).requires_grad_(True)
            if ref is not None:
                ref_mod, kind = ref[0].ClipLoss(local_loss=False, gather_with_grad=True, cache_labels=True, rank=rank, world_size=world), "reference (oracle/_ref)"
                ref_fn = lambda: ref_mod(A, B, 1.0)                                    # noqa: E731
            else:
                kind = "port"
                ref_fn = lambda: oc.clip_loss_port_distributed(A, B, 1.0, rank=rank, world_size=world, local_loss=False, gather_with_grad=True)   # noqa: E731
            ours = ClipLoss(local_loss=False, gather_with_grad=True, cache_labels=True, rank=rank, world_size=world, loss_dtype=torch.float32)

            def step(fn):
                A.grad = None; B.grad = None
                loss = fn()
                loss.backward()
                return loss

            def timed_world(fn):
                for _ in range(3):
                    step(fn)
                dist.barrier(); torch.cuda.synchronize()
                ts = []
                for _ in range(args.reps):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); loss = step(fn); e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                t = torch.tensor([sorted(ts)[len(ts) // 2]], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return float(t.item()), float(loss.detach().float())
            try:
                torch.cuda.reset_peak_memory_stats()
                t_e, l_e = timed_world(ref_fn)
                peak = torch.cuda.max_memory_allocated() / 2 ** 30
            except torch.cuda.OutOfMemoryError:
                t_e, l_e, peak = None, None, None
            torch.cuda.reset_peak_memory_stats()
            t_o, l_o = timed_world(lambda: ours(A, B))
            peak_o = torch.cuda.max_memory_allocated() / 2 ** 30
            row = dict(world=world, N=N, n=n, d=args.dim, dtype=str(dtype).split(".")[-1], tf32=tf32, eager_impl=kind, eager_ms=t_e,
                       fused_ms=t_o, speedup=(t_e / t_o) if t_e else None, eager_peak_gib=peak, fused_peak_gib=peak_o, eager_loss=l_e,
                       fused_loss=l_o, eager_samples_per_s=(N / (t_e * 1e-3)) if t_e else None, fused_samples_per_s=N / (t_o * 1e-3))
            if rank == 0:
                print(json.dumps(row), flush=True)
                rows.append(row)
            del A, B
            torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
