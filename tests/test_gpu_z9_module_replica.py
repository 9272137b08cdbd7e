"""GPU: the drop-ins (ClipLoss, BaseEncoder heads, RetrievalMetric) replay what the reference's own
OneProtLitModule recorded for its training_step (L1 term, gradient clipping, SGD), validation_step
(RetrievalMetric) and test_step (tensor logit_scale on already scaled features - the two-reference path)
in tests/golden/module_steps.npz.  Not yet run on hardware."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,loss_rtol,param_cos,metric_tol", [(torch.float32, 2e-3, 0.9999, 0.09), (torch.bfloat16, 5e-2, 0.995, 0.17)])
def test_drop_ins_reproduce_the_reference_modules_own_steps(dtype, loss_rtol, param_cos, metric_tol):
    from tests.module_replica import compare, run_replica
    g, out = run_replica(dtype, device="cuda")
    compare(g, out, loss_rtol=loss_rtol, param_cos=param_cos, metric_tol=metric_tol)
