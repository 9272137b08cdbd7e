"""Runs one tensor-core kernel of the path in isolation (for ncu captures / quick timing).

    python tools/run_kernel.py {fwd|fwd_e|dz|dz_e|da|db} [rows] [N] [d] [reps]

fwd_e / dz_e: the stored-exponentials pair (forward that keeps e_ij, in-place rescale; dz_e reports GB/s).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oneprot_b200 import kernels as K
from tools import synthetic as oc

which = sys.argv[1] if len(sys.argv) > 1 else "dz"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
N = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
d = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
dev = torch.device("cuda")
a, b = oc.synthetic_pair(N, d, seed=1)
A, B = a.to(dev), b.to(dev)
scale = torch.ones(1, device=dev)
stats = torch.zeros(4, device=dev)
diag = torch.empty(N, device=dev)
K.rowstats(A, B, 0, diag, stats)
rowsum = torch.empty(rows, device=dev); colsum = torch.empty(N, device=dev)
ldw = (N + 63) // 64 * 64
Wz = torch.zeros(rows, ldw, dtype=torch.bfloat16, device=dev)
wr = torch.full((rows,), 1e-6, device=dev); wc = torch.full((N,), 1e-6, device=dev); dg = torch.full((rows,), 1e-3, device=dev)
dA = torch.empty(rows, d, dtype=torch.bfloat16, device=dev); dB = torch.empty(N, d, dtype=torch.bfloat16, device=dev)
scratch = None
fns = {
    "fwd": lambda: K.fwd_sums(A[:rows], B, scale, stats, rowsum, colsum, scratch),
    "fwd_e": lambda: K.fwd_sums(A[:rows], B, scale, stats, rowsum, colsum, scratch, keep=Wz),
    "dz": lambda: K.dz_panel(A[:rows], B, 0, scale, stats, wr, wc, dg, Wz),
    "dz_e": lambda: K.dz_from_exp(Wz, rows, N, 0, wr, wc, dg),
    "da": lambda: K.gemm_bf16(Wz, False, B, True, rows, d, N, out=dA),
    "db": lambda: K.gemm_bf16(Wz, True, A[:rows], True, N, d, rows, out=dB),
}
fn = fns[which]
if which in ('da', 'db'):
    fns['dz'](); torch.cuda.synchronize()   # realistic panel contents (zeros would lower the MMA power)
if which == "dz_e":
    fns["fwd_e"](); torch.cuda.synchronize()   # real exponentials in the panel
if which in ("fwd", "fwd_e"):
    scratch = torch.empty(K.fwd_scratch_bytes(rows, N), dtype=torch.uint8, device=dev)
fn(); torch.cuda.synchronize()
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
if which == "dz_e":
    by = 4.0 * rows * N
    print(f"{which}: rows={rows} N={N}  ms={min(ts):.4f} (min of {reps}; all {[round(t,4) for t in ts]})  GB/s={by/min(ts)/1e6:.0f}")
    sys.exit(0)
fl = 2.0 * rows * N * d
print(f"{which}: rows={rows} N={N} d={d}  ms={min(ts):.4f} (min of {reps}; all {[round(t,4) for t in ts]})  TFLOP/s={fl/min(ts)/1e9:.1f}")
