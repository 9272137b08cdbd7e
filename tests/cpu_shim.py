"""Optional CPU shim for the GPU test files (ONEPROT_CPU_SHIM=1 python -m pytest tests/test_gpu_z3_heads.py -m gpu ...).

Purpose: shake Python-level mistakes (shapes, argument order, API misuse, impossible tolerances)
out of GPU tests BEFORE they cost GPU minutes.  Every "cuda" device becomes the CPU, streams /
events are stubs and the kernel provider is the float64 emulation of tests/fake_kernels.py, so a pass
here says nothing about the CUDA kernels - only that the test and the host logic around them are
sound.  Never active unless the environment variable is set; the product is not touched."""
import contextlib
import functools
import time

import torch


def _map_dev(d):
    if d is None:
        return None
    if isinstance(d, str):
        return "cpu" if d.startswith("cuda") else d
    if isinstance(d, int):
        return "cpu"
    return torch.device("cpu") if getattr(d, "type", None) == "cuda" else d


def _wrap_factory(fn):
    @functools.wraps(fn)
    def w(*a, **k):
        if "device" in k:
            k["device"] = _map_dev(k["device"])
        return fn(*a, **k)
    return w


class _Event:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max(1e-3, (other.t - self.t) * 1e3)


class _Stream:
    cuda_stream = 0x1000

    def __init__(self, device=None):
        pass

    def record_event(self):
        return _Event()

    def wait_event(self, e):
        pass

    def wait_stream(self, s):
        pass


def install(provider=None):
    """provider "fake" (default): float64 emulation of the kernel contracts; "emu" (ONEPROT_CPU_SHIM=emu):
    the library's own kernel source under the CPU emulation of tests/emu - slow, but the real kernels."""
    import os
    provider = provider or os.environ.get("ONEPROT_CPU_SHIM")
    if provider == "lib":
        # the WHOLE library compiled for the CPU (tests/emu/build_full_lib.py): the product's own kernels.py
        # wrappers, ctypes signatures, C host functions (incl. the step sequencer) and kernel source all run
        from tests.emu import build_full_lib
        os.environ["ONEPROT_LIB"] = build_full_lib.build()
        from oneprot_b200 import _lib, kernels
        _lib.LIB_PATH = os.environ["ONEPROT_LIB"]
        kernels._DRY = lambda: 0            # accept CPU tensors; "current stream" handle 0
        _install_torch_stubs()
        return
    from oneprot_b200 import clip_loss, epilogue, heads, module_steps, retrieval
    if provider == "emu":
        from tests import emu_kernels as fake_kernels
    else:
        from tests import fake_kernels
    _install_torch_stubs()
    for mod in (clip_loss, epilogue, heads, module_steps, retrieval):
        mod._KERNELS = fake_kernels


def _install_torch_stubs():
    for name in ("empty", "zeros", "ones", "full", "randn", "rand", "randint", "tensor", "arange", "eye", "empty_like", "zeros_like"):
        setattr(torch, name, _wrap_factory(getattr(torch, name)))
    torch.Tensor.cuda = lambda self, *a, **k: self.clone()        # a new tensor, like a real host-to-device copy
    torch.nn.Module.cuda = lambda self, *a, **k: self
    _to = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple(_map_dev(x) if isinstance(x, (str, torch.device)) else x for x in a)
        if "device" in k:
            k["device"] = _map_dev(k["device"])
        return _to(self, *a, **k)
    torch.Tensor.to = to
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    cur = _Stream()
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.current_stream = lambda *a, **k: cur
    torch.cuda.Stream = _Stream
    torch.cuda.Event = _Event
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.device_count = lambda: 1
    torch.cuda.max_memory_allocated = lambda *a, **k: 0
    torch.cuda.reset_peak_memory_stats = lambda *a, **k: None
    torch.cuda.empty_cache = lambda: None
