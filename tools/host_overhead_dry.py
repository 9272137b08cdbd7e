"""Host-side cost of one ClipLoss fwd+bwd WITHOUT a GPU: both hosts (kernel-by-kernel Python path and
the C step sequencer) run in the library's dry launch-trace mode on CPU tensors, so what is timed is
the Python / ctypes / torch-allocator work per step - not the CUDA launch calls themselves (~2-4 us
each on a B200 host, the same number of launches on both paths).

    python tools/host_overhead_dry.py [n] [d]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oneprot_b200 import clip_loss as cl  # noqa: E402
from oneprot_b200 import comm as comm_mod  # noqa: E402
from oneprot_b200 import kernels as K  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    A = torch.randn(n, d).to(torch.bfloat16)
    B = torch.randn(n, d).to(torch.bfloat16)
    scale = torch.ones(1)
    # Wz panels are only address labels here: bound them so that the CPU allocations stay small
    pb = 2 * ((n + 63) // 64 * 64) * 128 * ((n + 127) // 128)
    for seq in (False, True):
        cfg = dict(world_size=1, rank=0, group=None, local_loss=False, gather_with_grad=False, loss_dtype=torch.float32,
                   panel_bytes=pb, host_sequencer=seq)
        local = comm_mod.LocalComm(K)
        cl._get_comm = lambda *a, **k: local
        ts = []
        with K.launch_trace(dry_stream=lambda: 0x1000) as tr:
            for it in range(220):
                a = A.requires_grad_(True)
                b = B.requires_grad_(True)
                t0 = time.perf_counter()
                loss, _, _ = cl._ClipLossFunction.apply(a, b, scale, cfg)
                loss.backward()
                ts.append(time.perf_counter() - t0)
                a.grad = b.grad = None
        ts = sorted(ts[20:])
        launches = sum(1 for ln in tr.lines if not ln.startswith(("memset", "copy", " "))) / 220
        print(f"{'C sequencer ' if seq else 'python host '} n={n} d={d}: median {1e6 * ts[len(ts) // 2]:7.1f} us/step, "
              f"p90 {1e6 * ts[int(len(ts) * 0.9)]:7.1f} us  ({launches:.0f} entry-point calls per step, tracing on)")


if __name__ == "__main__":
    main()
