"""Runs a GPU tool / script on the CPU stand-ins of tests/cpu_shim.py to shake out Python-level mistakes
before it costs GPU minutes:   python tests/run_tool_cpu_shim.py tools/bench_heads.py 512
(test hygiene only: the kernels are the float64 emulation, the numbers mean nothing)."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import cpu_shim, fake_kernels  # noqa: E402

cpu_shim.install()
from oneprot_b200 import kernels  # noqa: E402

for name in dir(fake_kernels):
    if not name.startswith("_") and callable(getattr(fake_kernels, name)) and hasattr(kernels, name):
        setattr(kernels, name, getattr(fake_kernels, name))
script = sys.argv[1]
sys.argv = sys.argv[1:]
runpy.run_path(script, run_name="__main__")
