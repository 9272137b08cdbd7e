"""Kernel timeline of ClipLoss steps on rank 0 (torch.profiler / CUPTI): per-kernel durations and idle gaps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from oneprot_b200 import ClipLoss
from tools import synthetic as oc
world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1: dist.init_process_group("nccl", device_id=dev)
N = int(os.environ.get("ONEPROT_BENCH_N", 32768)); n = N // world
a, b = oc.synthetic_pair(n, 1024, seed=1234, rank=rank)
A = a.to(dev).requires_grad_(True); B = b.to(dev).requires_grad_(True)
m = ClipLoss(local_loss=False, gather_with_grad=True, rank=rank, world_size=world)
def step():
    A.grad = None; B.grad = None
    m(A, B).backward()
for _ in range(8): step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
acts = [ProfilerActivity.CUDA, ProfilerActivity.CPU]
with profile(activities=acts) as prof:
    for _ in range(6): step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    # split into steps at mc_store / all_gather boundaries: just print the sequence of the 3rd step
    names = [e.name for e in evs]
    starts = [i for i, nme in enumerate(names) if "rowstats" in nme]
    s, e_ = starts[3], starts[4]
    prev_end = None
    tot_k = 0.0
    print(f"W={world} n={n}: one step, kernels in order (start offset us, duration us, gap before us)")
    for ev in evs[s:e_]:
        st, en = ev.time_range.start, ev.time_range.end
        gap = (st - prev_end) if prev_end is not None else 0.0
        print(f"  +{st - evs[s].time_range.start:8.1f}  dur {en - st:8.1f}  gap {gap:7.1f}  {ev.name[:70]}")
        prev_end = max(prev_end, en) if prev_end is not None else en
        tot_k += en - st
    print(f"step span {evs[e_].time_range.start - evs[s].time_range.start:.1f} us, sum of kernel durations {tot_k:.1f} us")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
