"""GPU: the dL/dZ panel kernel with L2 cache hints (ONEPROT_DZ_L2_HINTS=1: panel stores evict-first,
operand loads evict-last; clip_s_kernel<DZ_L2>) gives bit-identical gradients to the default kernel.
The knob is read once per process, so the hinted run is a subprocess.  Not yet run on hardware."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, torch
sys.path.insert(0, %r)
from oneprot_b200 import ClipLoss
from tools.synthetic import synthetic_pair
a, b = synthetic_pair(1536, 256, seed=9)
A = a.cuda().requires_grad_(True); B = b.cuda().requires_grad_(True)
keep = len(sys.argv) > 2 and sys.argv[2] == "keep"                                # stored-exponentials forward (clip_s_kernel<FWD_E[_L2]>)
loss = ClipLoss(loss_dtype=torch.float32, panel_bytes=2 * 1536 * 640, keep_exp=keep)(A, B)     # three panels
loss.backward(); torch.cuda.synchronize()
torch.save({"loss": loss.detach().cpu(), "dA": A.grad.cpu(), "dB": B.grad.cpu()}, sys.argv[1])
"""


def _run(tmp_path, name, env_extra, *extra):
    out = str(tmp_path / name)
    env = dict(os.environ, **env_extra)
    env.pop("ONEPROT_DZ_L2_HINTS", None) if not env_extra else None
    p = subprocess.run([sys.executable, "-c", SCRIPT % ROOT, out, *extra], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return torch.load(out)


def test_l2_hinted_panel_kernel_is_bit_identical(tmp_path):
    base = _run(tmp_path, "base.pt", {})
    hint = _run(tmp_path, "hint.pt", {"ONEPROT_DZ_L2_HINTS": "1"})
    assert base["loss"].item() == hint["loss"].item()
    assert torch.equal(base["dA"], hint["dA"]) and torch.equal(base["dB"], hint["dB"])


def test_l2_hinted_keeping_forward_is_bit_identical(tmp_path):
    base = _run(tmp_path, "kbase.pt", {}, "keep")
    hint = _run(tmp_path, "khint.pt", {"ONEPROT_DZ_L2_HINTS": "1"}, "keep")
    assert base["loss"].item() == hint["loss"].item()
    assert torch.equal(base["dA"], hint["dA"]) and torch.equal(base["dB"], hint["dB"])
