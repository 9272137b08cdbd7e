"""CPU: the per-step multi-modality driver (oneprot_b200/module_steps.py) - the fused L1 term against
torch.abs(x).mean() and its autograd, on the float64 stand-ins and on the kernel source under the CPU
emulation; the step logic of OneProtLitModule (modality schedule, seqsim handling, optimizer steps between
the pairs).  The full trajectory against the reference's own module is in tests/test_heads_cpu.py /
tests/test_emulated_product_cpu.py (tests/golden/module_steps.npz)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from tests import fake_kernels


def _provider(name):
    if name == "fake":
        return fake_kernels
    from tests import emu_kernels
    if not emu_kernels.available():
        pytest.skip("g++ or the CUDA headers are not available")
    emu_kernels.prebuild()
    return emu_kernels


@pytest.fixture(params=["fake", "emu"])
def ms(request, monkeypatch):
    from oneprot_b200 import module_steps
    monkeypatch.setattr(module_steps, "_KERNELS", _provider(request.param))
    return module_steps


@pytest.mark.parametrize("shape", [(12, 32), (7, 13), (3, 5, 24), (300, 1024)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_mean_abs_matches_torch(ms, shape, dtype):
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g)
    x.view(-1)[::7] = 0.0                                    # sign(0) = 0, like torch.abs' gradient
    x = x.to(dtype).requires_grad_(True)
    y = ms.mean_abs(x)
    assert y.dim() == 0 and y.dtype == dtype
    want = x.detach().double().abs().mean()
    assert abs(float(y.detach()) - float(want)) <= (1e-6 if dtype == torch.float32 else 4e-3) * float(want)
    (y * 3.0).backward()
    ref = 3.0 * torch.sign(x.detach().double()) / x.numel()
    assert x.grad.shape == x.shape and x.grad.dtype == dtype
    assert np.allclose(x.grad.double().numpy(), ref.numpy(), rtol=1e-2 if dtype == torch.bfloat16 else 1e-6, atol=0)


def test_mean_abs_is_deterministic_and_rejects_empty(ms):
    x = torch.randn(513, 72)
    assert float(ms.mean_abs(x)) == float(ms.mean_abs(x.clone()))
    with pytest.raises(ValueError):
        ms.mean_abs(torch.zeros(0, 8))


class _Enc(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.w = nn.Parameter(torch.eye(d))
        self.norm = nn.Sequential(nn.Identity(), nn.Module())
        self.norm[1].log_logit_scale = torch.tensor(0.5)

    def forward(self, x):
        return x @ self.w


def test_step_logic_follows_the_reference_module(monkeypatch):
    """oneprot_module.py:80-108: 'struct_token' only before train_on_all_modalities_after_step, 'seqsim' dropped unless
    use_seqsim and encoded by the SEQUENCE encoder, one optimizer step per pair (the second pair sees updated weights)."""
    from oneprot_b200 import module_steps
    monkeypatch.setattr(module_steps, "_KERNELS", fake_kernels)
    d = 8
    net = nn.ModuleDict({k: _Enc(d) for k in ("sequence", "struct_token", "text")})
    seen = []

    def loss_fn(a, b, scale=1.0):
        seen.append((a.detach().clone(), b.detach().clone(), float(scale)))
        return ((a - b) ** 2).mean() * scale

    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    st = module_steps.ModalitySteps(net, loss_fn, opt, use_l1_regularization=True, train_on_all_modalities_after_step=1)
    g = torch.Generator().manual_seed(0)
    batch = {k: (torch.randn(4, d, generator=g), torch.randn(4, d, generator=g), None, None) for k in ("text", "struct_token", "seqsim")}
    w0 = net["sequence"].w.detach().clone()
    losses = st.training_step(batch)
    assert len(losses) == 1 and st.global_step == 1                      # warm-up phase: struct_token only
    assert torch.equal(seen[0][0], batch["struct_token"][0] @ w0)
    w1 = net["sequence"].w.detach().clone()
    assert not torch.equal(w0, w1)
    seen.clear()
    losses = st.training_step(batch)
    assert len(losses) == 2 and st.global_step == 3                      # text, struct_token; seqsim dropped
    assert torch.equal(seen[0][0], batch["text"][0] @ w1)
    assert not torch.equal(seen[1][0], batch["struct_token"][0] @ w1)    # encoded AFTER the optimizer step of the first pair
    st.use_seqsim = True
    seen.clear()
    assert len(st.training_step(batch)) == 3
    # gradient clipping to norm 1.0 before every optimizer step (oneprot_module.py:105)
    total = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in net.parameters() if p.grad is not None))
    assert float(total) <= 1.0 + 1e-6
    # test_step passes the modality's exp(log_logit_scale) as logit_scale (oneprot_module.py:142)
    seen.clear()
    out = st.test_step({"text": batch["text"]})
    assert set(out) == {"text"} and abs(seen[0][2] - float(torch.tensor(0.5).exp())) < 1e-6
    # validation_step: (sequence_inputs, modality_inputs, modality, _) and the 'seqsim' -> sequence encoder rule
    seen.clear()
    st.validation_step((batch["seqsim"][0], batch["seqsim"][1], "seqsim", None))
    assert torch.equal(seen[0][1], batch["seqsim"][1] @ net["sequence"].w.detach())
