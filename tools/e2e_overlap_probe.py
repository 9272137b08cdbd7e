"""Why does the prefetched H2D copy not hide under the step at 8 GPUs?  (torchrun, one rank per GPU)
 A  device-resident steps alone
 B  the same steps while a copy stream copies the pair in a free-running loop (no dependency on the steps)
 C  PinnedPairPrefetcher, 2 slots, submit one ahead (bench.py's pipelined leg)
 D  PinnedPairPrefetcher, 3 slots, submit two ahead
Prints ms per step (max over ranks) and, for B, the mean duration of one pair copy under load."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from oneprot_b200 import ClipLoss
from oneprot_b200.prefetch import PinnedPairPrefetcher
from tools.synthetic import synthetic_global_rows
world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
if world > 1: dist.init_process_group("nccl", device_id=dev)
N, d = int(os.environ.get("PROBE_N", 32768)), 1024; n = N // world
a, b = synthetic_global_rows(rank * n, n, d)
a_pin, b_pin = a.pin_memory(), b.pin_memory()
A = a.to(dev).requires_grad_(True); B = b.to(dev).requires_grad_(True)
m = ClipLoss(local_loss=False, gather_with_grad=True, rank=rank, world_size=world, graph=bool(os.environ.get("PROBE_GRAPH")))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
K = 60
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
def reduce_max(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
def timed(fn, k=K, do_flush=True):
    for _ in range(8): fn()
    barrier(); evs = []
    for _ in range(k):
        if do_flush: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    r = reduce_max(sum(x.elapsed_time(y) for x, y in evs) / k); barrier(); return r
def step_dev():
    A.grad = None; B.grad = None
    m(A, B).backward()
out = {"world": world, "A_device_ms": timed(step_dev)}
# B: free-running copies
cs = torch.cuda.Stream(device=dev); bufA, bufB = torch.empty_like(A), torch.empty_like(B); cev = []
def step_with_free_copy():
    with torch.cuda.stream(cs):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); bufA.copy_(a_pin, non_blocking=True); bufB.copy_(b_pin, non_blocking=True); e1.record(); cev.append((e0, e1))
    step_dev()
out["B_steps_with_free_running_copies_ms"] = timed(step_with_free_copy)
torch.cuda.synchronize()
out["B_copy_under_load_ms"] = reduce_max(sum(x.elapsed_time(y) for x, y in cev[-K:]) / K)
for tag, slots in (("C_prefetch_2slots_ms", 2), ("D_prefetch_3slots_ms", 3)):
    pf = PinnedPairPrefetcher(dev, slots=slots)
    for _ in range(slots - 1): pf.submit(a_pin, b_pin)
    def step_pf():
        Ad, Bd = pf.next(); pf.submit(a_pin, b_pin)
        Ad.requires_grad_(True); Bd.requires_grad_(True)
        m(Ad, Bd).backward()
    out[tag] = timed(step_pf)
    out[tag.replace("_ms", "_noflush_ms")] = timed(step_pf, do_flush=False)
out["A_device_noflush_ms"] = timed(step_dev, do_flush=False)
if rank == 0: print(out, flush=True)
if world > 1: dist.barrier(); dist.destroy_process_group()
