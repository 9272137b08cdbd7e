"""GPU: the per-step driver and the drop-ins (ModalitySteps, ClipLoss, BaseEncoder heads, RetrievalMetric) replay what the reference's own
OneProtLitModule recorded for its training_step (L1 term, gradient clipping, SGD), validation_step
(RetrievalMetric) and test_step (tensor logit_scale on already scaled features - the two-reference path)
in tests/golden/module_steps.npz.  Green on B200 since round 2."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,loss_rtol,param_cos,metric_tol", [(torch.float32, 2e-3, 0.9999, 0.09), (torch.bfloat16, 5e-2, 0.995, 0.17)])
def test_drop_ins_reproduce_the_reference_modules_own_steps(dtype, loss_rtol, param_cos, metric_tol):
    from tests.module_replica import compare, run_replica
    g, out = run_replica(dtype, device="cuda")
    compare(g, out, loss_rtol=loss_rtol, param_cos=param_cos, metric_tol=metric_tol)


@pytest.mark.parametrize("shape", [(12, 32), (7, 13), (4096, 1024)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_l1_term_matches_torch(shape, dtype):
    """mean_abs (abs_sum_kernel / abs_mean_bwd_kernel) == torch.abs(x).mean() and its gradient (oneprot_module.py:43-44)."""
    from oneprot_b200 import mean_abs
    g = torch.Generator().manual_seed(sum(shape))
    xh = torch.randn(shape, generator=g)
    xh.view(-1)[::7] = 0.0
    x = xh.to(dtype).cuda().requires_grad_(True)
    y = mean_abs(x)
    want = x.detach().double().abs().mean()
    assert abs(float(y.detach()) - float(want)) <= (3e-6 if dtype == torch.float32 else 4e-3) * float(want)
    (y * 3.0).backward()
    ref = 3.0 * torch.sign(x.detach().double()) / x.numel()
    assert torch.allclose(x.grad.double(), ref, rtol=1e-2 if dtype == torch.bfloat16 else 1e-6, atol=0)
    assert float(mean_abs(x.detach())) == float(mean_abs(x.detach().clone()))       # deterministic
