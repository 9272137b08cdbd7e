"""Host-side enqueue cost of one ClipLoss fwd+bwd (no device sync inside the loop) and the wall time
per step at sizes where the device work is small, for the three hosts:

    python tools/host_overhead.py [n] [d]
      python   kernel-by-kernel Python host (default path)
      seq      host-side step sequencer (ClipLoss(host_sequencer=True), csrc/clip_sequence.cu)
      graph    CUDA-graph replay (ClipLoss(graph=True), oneprot_b200/graphed.py)
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oneprot_b200 import ClipLoss  # noqa: E402
from tools import synthetic as oc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
a, b = oc.synthetic_pair(n, d, seed=1)
A = a.cuda().requires_grad_(True)
B = b.cuda().requires_grad_(True)
ref = None
for name, kw in (("python", dict(host_sequencer=False)), ("seq", dict(host_sequencer=True)), ("graph", dict(graph=True))):
    try:
        m = ClipLoss(loss_dtype=torch.float32, **kw)

        def step():
            A.grad = None; B.grad = None
            loss = m(A, B)
            loss.backward()
            return loss

        for _ in range(10):
            loss = step()
        torch.cuda.synchronize()
        val = (loss.item(), A.grad.clone(), B.grad.clone())
        if ref is None:
            ref = val
        same = val[0] == ref[0] and torch.equal(val[1], ref[1]) and torch.equal(val[2], ref[2])
        t0 = time.perf_counter()
        for _ in range(300):
            step()
        t_enq = (time.perf_counter() - t0) / 300
        torch.cuda.synchronize()
        t_tot = (time.perf_counter() - t0) / 300
        print(f"{name:7s} n={n} d={d}: host enqueue {t_enq * 1e6:7.1f} us/step, wall incl. device {t_tot * 1e6:7.1f} us/step, "
              f"bit-identical to the python host: {same}", flush=True)
    except Exception as e:   # keep going: the other hosts are still worth measuring
        print(f"{name:7s} FAILED: {e!r}", flush=True)
