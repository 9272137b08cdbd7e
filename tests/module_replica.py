"""The reference's manual-optimisation steps (oneprot_module.py:80-146: training_step with the L1 term,
validation_step with RetrievalMetric, test_step with the tensor logit_scale) run through this package's
driver and drop-ins (ModalitySteps, ClipLoss, BaseEncoder, RetrievalMetric) on the inputs and initial
parameters of tests/golden/module_steps.npz - the fixture recorded from the reference's own
OneProtLitModule (oracle/make_golden.py::module_cases).  Shared by the CPU and GPU tests."""
import ast

import numpy as np
import torch
import torch.nn as nn

from tests.helpers import bf16_from_bits, load_golden


def run_replica(dtype, device="cpu"):
    from oneprot_b200 import BaseEncoder, ClipLoss, ModalitySteps, RetrievalMetric
    g = load_golden("module_steps.npz")
    spec = ast.literal_eval(g["spec"].item().decode())
    net = nn.ModuleDict({k: BaseEncoder(dm, 32, proj_type=pt, use_logit_scale=uls, learnable_logit_scale=False, pooling_type=pool)
                         for k, (dm, pt, uls, pool) in spec.items()})
    net.load_state_dict({k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("init:")}, strict=True)
    net = net.to(device=device, dtype=dtype)
    loss_fn = ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=0, world_size=1)     # oneprot_module.py:48-56
    opt = torch.optim.SGD(net.parameters(), lr=float(g["lr"]), momentum=float(g["momentum"]))

    def batch(tag, mod):
        seq = bf16_from_bits(g[f"{tag}:{mod}:seq_bf16"]).reshape(tuple(g[f"{tag}:{mod}:seq_shape"]))
        x = bf16_from_bits(g[f"{tag}:{mod}:mod_bf16"]).reshape(tuple(g[f"{tag}:{mod}:mod_shape"]))
        return seq.to(device=device, dtype=dtype), x.to(device=device, dtype=dtype)

    mods = ("text", "struct_graph")
    metrics = {f"val_{m}": RetrievalMetric() for m in mods}
    steps = ModalitySteps(net, loss_fn, opt, use_l1_regularization=True, metrics=metrics)       # oneprot_module.py:10-41
    out = {"train": [], "val": [], "test": [], "valmetric": {}}
    for s in range(len(g["train_losses"]) // 2):
        b = {}
        for mod in mods:
            seq, x = batch(f"train{s}", mod)
            b[mod] = (seq, x, None, None)
        out["train"] += [float(v) for v in steps.training_step(b)]               # :80-108
    vb = {}
    for mod in mods:
        seq, x = batch("val", mod)
        vb[mod] = (seq, x, None, None)
        out["val"].append(float(steps.validation_step((seq, x, mod, None))))     # :110-121
        out["valmetric"][mod] = metrics["val_" + mod].compute()
    out["test"] = [float(v) for v in steps.test_step(vb).values()]               # :137-146
    assert steps.global_step == len(g["train_losses"])
    out["final"] = {k: v.detach().double().cpu().numpy() for k, v in net.state_dict().items()}
    return g, out


def compare(g, out, loss_rtol, param_cos, metric_tol):
    from tests.helpers import cosine
    for key in ("train", "val", "test"):
        want = g[f"{key}_losses"]
        assert len(out[key]) == len(want)
        for i, (a, b) in enumerate(zip(out[key], want)):
            assert abs(a - b) <= loss_rtol * abs(b), (key, i, a, b)
    for k, v in g.items():
        if k.startswith("final:") and np.ndim(v) >= 1:
            assert cosine(out["final"][k[6:]], v) >= param_cos, k
        if k.startswith("valmetric:"):
            _, mod, name = k.split(":", 2)
            tol = metric_tol * (12 if "median" in name else 1)
            assert abs(float(out["valmetric"][mod][name]) - float(v)) <= tol, (k, out["valmetric"][mod][name], float(v))
