"""Runs bench.py's GPU arm (run_ours, world 1) on the CPU with stand-ins, to exercise its CONTROL FLOW -
line construction, serial and pipelined e2e legs, fallback, watchdog, single JSON line - where no GPU
exists.  Test infrastructure (started by tests/test_bench_cpu.py in a subprocess): CUDA streams /
events / pinned memory are stubs, the kernels are the float64 emulation of tests/fake_kernels.py and
the problem is tiny (ONEPROT_BENCH_N / _D).  The numbers it prints mean nothing.

    python tests/bench_cpu_harness.py [pipelined|fallback|m5|lib-pipelined|lib-fallback]

The lib-* modes replace the float64 stand-ins by the library itself compiled for the CPU (tests/emu), so
bench.py's calls go through the real kernels.py wrappers, ctypes signatures and C entry points.
"""
import contextlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("ONEPROT_BENCH_N", "128" if len(sys.argv) > 1 and sys.argv[1].startswith("lib") else "256")
os.environ.setdefault("ONEPROT_BENCH_D", "64")

import torch  # noqa: E402
import torch.distributed.nn  # noqa: E402,F401  (imported by the reference's loss.py; must precede the torch.device stand-in)

from tests import fake_kernels  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fallback"
cpu = torch.device("cpu")
real_device = torch.device


class FakeEvent:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max(1e-3, (other.t - self.t) * 1e3)


class FakeStream:
    cuda_stream = 0x1000

    def __init__(self, device=None):
        pass

    def record_event(self):
        e = FakeEvent()
        e.record()
        return e

    def wait_event(self, e):
        pass

    def wait_stream(self, s):
        pass


_cur = FakeStream()
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.Event = FakeEvent
torch.cuda.Stream = FakeStream
torch.cuda.current_stream = lambda *a, **k: _cur
torch.cuda.stream = lambda s: contextlib.nullcontext()
torch.cuda.device = lambda d: contextlib.nullcontext()
torch.Tensor.pin_memory = lambda self, *a, **k: self
torch.Tensor.is_pinned = lambda self, *a, **k: True


class _Dev:
    """torch.device('cuda', i) -> the CPU device (everything in bench.py goes through torch.device)."""
    def __call__(self, *a, **k):
        if a and (a[0] == "cuda" or getattr(a[0], "type", None) == "cuda"):
            return cpu
        return real_device(*a, **k)


torch.device = _Dev()

from oneprot_b200 import clip_loss, kernels  # noqa: E402

if mode.startswith("lib"):
    # the product's own wrappers and C host code over the library compiled for the CPU (tests/emu/build_full_lib.py)
    from tests.emu import build_full_lib
    from oneprot_b200 import _lib
    os.environ["ONEPROT_LIB"] = _lib.LIB_PATH = build_full_lib.build()
    kernels._DRY = lambda: 0
    mode = mode[len("lib-"):] or "fallback"
else:
    clip_loss._KERNELS = fake_kernels
    for name in ("rowstats", "fwd_sums", "dz_panel", "dz_from_exp", "gemm_bf16", "loss_finalize", "bwd_weights"):
        setattr(kernels, name, getattr(fake_kernels, name))

if mode == "pipelined":
    # let the prefetcher believe it stages into CUDA memory
    from oneprot_b200 import prefetch

    class _D:
        type = "cuda"
    prefetch.torch = type("T", (), {"device": staticmethod(lambda d: _D()), "cuda": torch.cuda, "Tensor": torch.Tensor,
                                    "empty": staticmethod(lambda *a, device=None, **k: torch.empty(*a, **k))})

import bench  # noqa: E402
import torch.distributed as dist  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", 1))
if world > 1:      # launched by torch.distributed.run: NCCL -> gloo
    _init = dist.init_process_group
    dist.init_process_group = lambda backend=None, device_id=None, **k: _init("gloo", **k)
sys.argv = ["bench.py", "--gpus", str(world), "--steps", "2" if "ONEPROT_LIB" in os.environ else "3", "--warmup", "3"]
if mode == "m5":      # BASELINE cfg 3: five pairs in sequence (tiny here), host variants that cannot run on the CPU are skipped
    sys.argv += ["--config", "modalities5", "--no-cpu-baseline"]
bench.main()
