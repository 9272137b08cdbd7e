"""Host-side enqueue cost of one ClipLoss fwd+bwd (no device sync inside the loop) and the
GPU-side time at a size where the device work is negligible."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oneprot_b200 import ClipLoss
from tools import synthetic as oc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
a, b = oc.synthetic_pair(n, 1024, seed=1)
A = a.cuda().requires_grad_(True); B = b.cuda().requires_grad_(True)
m = ClipLoss()
def step():
    A.grad = None; B.grad = None
    m(A, B).backward()
for _ in range(10): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
t_enq = (time.perf_counter() - t0) / 200
torch.cuda.synchronize()
t_tot = (time.perf_counter() - t0) / 200
print(f"n={n}: host enqueue {t_enq*1e6:.0f} us/step, wall incl. device {t_tot*1e6:.0f} us/step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(100): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
