"""GPU: the drop-in modules behind a replica of OneProt's manual-optimisation training step
(reference: src/models/oneprot_module.py:80-108): per modality  features -> ClipLoss (+ the L1
term of :99-101) -> backward -> clip_grad_norm(1.0) -> optimizer step.  The same loop is run with
the oracle's torch port of the reference ops in fp32; loss trajectories and final weights must
agree.  Encoders are stand-in Linear projections (the real towers are out of scope)."""
import copy

import pytest
import torch
import torch.nn as nn

from oracle import clip_oracle as oc

pytestmark = pytest.mark.gpu


class _RefNormalize(nn.Module):          # base_encoder.py:6-12 restated with the library op
    def forward(self, x):
        return torch.nn.functional.normalize(x, dim=-1, p=2)


class _RefScale(nn.Module):              # base_encoder.py:15-33, non-learnable as in the shipped configs
    def __init__(self, s=1 / 0.07, mx=100.0):
        super().__init__()
        self.register_buffer("log_logit_scale", torch.log(torch.tensor(s)))
        self.mx = mx

    def forward(self, x):
        return torch.clip(self.log_logit_scale.exp(), max=self.mx) * x


def _towers(ours: bool, dtype):
    from oneprot_b200 import LearnableLogitScaling, Normalize
    torch.manual_seed(0)
    seq = nn.Linear(96, 64, bias=False)
    mod = nn.Linear(80, 64, bias=False)
    if ours:
        seq_norm = nn.Sequential(Normalize(dim=-1))
        mod_norm = nn.Sequential(Normalize(dim=-1), LearnableLogitScaling(learnable=False))
    else:
        seq_norm = nn.Sequential(_RefNormalize())
        mod_norm = nn.Sequential(_RefNormalize(), _RefScale())
    net = nn.ModuleDict({"seq": seq, "mod": mod, "seq_norm": seq_norm, "mod_norm": mod_norm}).cuda().to(dtype)
    return net


def _run(ours: bool, dtype, steps=6):
    from oneprot_b200 import ClipLoss
    net = _towers(ours, dtype)
    # SGD keeps the comparison linear in the gradients (Adam's sign-like update amplifies the
    # bf16-panel noise of near-zero gradient entries into lr-sized weight differences)
    opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr=0.5, momentum=0.9)
    w0 = {k: v.detach().float().cpu().clone() for k, v in net.state_dict().items()}
    loss_fn = ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=0, world_size=1) if ours else None
    g = torch.Generator().manual_seed(5)
    losses = []
    for _ in range(steps):
        xs = torch.randn(200, 96, generator=g).cuda().to(dtype)
        xm = (xs[:, :80] + 0.3 * torch.randn(200, 80, generator=g).cuda().to(dtype))
        seq_f = net["seq_norm"](net["seq"](xs))
        mod_f = net["mod_norm"](net["mod"](xm))
        opt.zero_grad()
        loss = loss_fn(seq_f, mod_f) if ours else oc.clip_loss_port(seq_f, mod_f, 1.0)
        loss = loss + 0.01 * (torch.abs(seq_f).mean() + torch.abs(mod_f).mean())      # oneprot_module.py:99-101
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        losses.append(float(loss.detach()))
    return losses, {k: v.detach().float().cpu() - w0[k] for k, v in net.state_dict().items()}


def test_training_step_fp32_matches_reference_ops():
    torch.backends.cuda.matmul.allow_tf32 = False
    ours, w_ours = _run(True, torch.float32)
    ref, w_ref = _run(False, torch.float32)
    for a, b in zip(ours, ref):
        assert abs(a - b) < 2e-4 * abs(b), (ours, ref)
    assert ours[-1] < ours[0]                      # it trains
    for k in ("seq.weight", "mod.weight"):      # accumulated weight updates point the same way
        a, b = w_ours[k].flatten().double(), w_ref[k].flatten().double()
        assert float(a @ b / (a.norm() * b.norm())) > 0.9995, k
        assert abs(float(a.norm() / b.norm()) - 1) < 1e-2, k
    # same state_dict keys as the reference modules: checkpoints stay loadable (train.py:73-82)
    assert set(w_ours) == set(w_ref)


def test_training_step_bf16_tracks_fp32_reference():
    ours, _ = _run(True, torch.bfloat16)
    ref, _ = _run(False, torch.float32)
    for a, b in zip(ours, ref):
        assert abs(a - b) < 3e-2 * abs(b), (ours, ref)
