"""Per-phase CUDA-event timing of one ClipLoss fwd+bwd (run under torchrun for W > 1).

Instruments the same call sequence as oneprot_b200/clip_loss.py by wrapping the kernel provider
and torch.distributed collectives with events (diagnostic tool, not part of the product path)."""
import os
import sys
import collections

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from oneprot_b200 import ClipLoss, clip_loss, kernels
from tools import synthetic as oc

world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = int(os.environ.get("ONEPROT_BENCH_N", 32768)); d = 1024; n = N // world
a, b = oc.synthetic_pair(n, d, seed=1234, rank=rank)
A = a.to(dev).requires_grad_(True); B = b.to(dev).requires_grad_(True)

events = []
def wrap(obj, name, label=None):
    fn = getattr(obj, name)
    def w(*args, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*args, **kw); e1.record()
        events.append((label or name, e0, e1)); return r
    setattr(obj, name, w)

class KW:  # proxy of the kernels module with timed entry points
    pass
kw = KW()
for name in dir(kernels):
    setattr(kw, name, getattr(kernels, name))
for name in ("rowstats", "fwd_sums", "loss_finalize", "bwd_weights", "dz_panel", "gemm_bf16", "sum_f32", "rowdot_bf16"):
    wrap(kw, name)
clip_loss._KERNELS = kw
if world > 1:
    for name in ("all_gather_into_tensor", "all_reduce", "reduce_scatter_tensor"):
        wrap(clip_loss.dist, name, "nccl." + name)

m = ClipLoss(local_loss=False, gather_with_grad=True, rank=rank, world_size=world)
def step():
    A.grad = None; B.grad = None
    m(A, B).backward()
for _ in range(5):
    step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
agg = collections.OrderedDict(); tot = []
for it in range(10):
    events.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record(); torch.cuda.synchronize()
    tot.append(e0.elapsed_time(e1))
    seen = collections.Counter()
    for label, a0, a1 in events:
        seen[label] += 1
        key = f"{label}#{seen[label]}"
        agg.setdefault(key, []).append(a0.elapsed_time(a1))
    # gaps: time not covered by any event
if rank == 0:
    print(f"W={world} n={n} total ms/step: mean {sum(tot)/len(tot):.3f} min {min(tot):.3f}")
    s = 0
    for k, v in agg.items():
        mv = sum(v) / len(v); s += mv
        print(f"  {k:38s} {mv*1e3:8.1f} us")
    print(f"  {'sum of phases':38s} {s*1e3:8.1f} us ; uncovered (gaps, torch ops) {(sum(tot)/len(tot)-s)*1e3:8.1f} us")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
