"""Per-step multi-modality driver: the step methods of the reference's ``OneProtLitModule``
(``src/models/oneprot_module.py:80-146``) without Lightning, over this package's drop-ins.

The reference trains with manual optimisation: for every modality of the batch it encodes the pair,
evaluates ``loss_fn`` (+ ``0.01 * (mean|seq| + mean|mod|)`` when ``use_l1_regularization``), clips
the gradient norm to 1.0 and steps the optimizer BEFORE the next modality (oneprot_module.py:92-108).
The optimizer step between the pairs makes the <= 5 problems of a batch sequential (pair m + 1 is
encoded with the parameters pair m just updated), so they are not grouped into one launch here
either; what the driver adds on B200 is that everything between the encoders' outputs and the
optimizer - loss, L1 term, their backward - is kernels of liboneprot_clip.so (``mean_abs`` below is
the fused L1 term: one read of the features forward, one sign pass backward), and that the loss may
replay as CUDA graphs / one C call per phase (``ClipLoss(graph=True)`` / ``host_sequencer=True``),
which is what removes the launch overhead that dominates at OneProt's batch sizes.

``validation_step`` / ``test_step`` keep the reference's semantics too, including test_step's use
of the modality's ``log_logit_scale.exp()`` as ``logit_scale`` on features that are already scaled
(oneprot_module.py:142).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import kernels as _cuda_kernels

_KERNELS = _cuda_kernels       # tests swap this for a stand-in that lives under tests/


class _MeanAbsFn(torch.autograd.Function):
    """``torch.abs(x).mean()`` (oneprot_module.py:43-44)."""

    @staticmethod
    def forward(ctx, x):
        K = _KERNELS
        x2 = x.detach().reshape(-1)
        if x2.dtype not in (torch.bfloat16, torch.float32):
            x2 = x2.float()
        count = x2.numel()
        pad = (-count) % 8
        if pad:
            x2 = torch.nn.functional.pad(x2, (0, pad))
        x2 = x2.contiguous()
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        K.abs_mean_fwd(x2, count, out)
        ctx.saved = (x2, count, x.shape, x.dtype)
        return out.reshape(()).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        K = _KERNELS
        x2, count, shape, dtype = ctx.saved
        g32 = g.detach().to(device=x2.device, dtype=torch.float32).reshape(1)
        gx = torch.empty_like(x2)
        K.abs_mean_bwd(x2, g32, count, gx)
        return gx[:count].reshape(shape).to(dtype)


def mean_abs(x: torch.Tensor) -> torch.Tensor:
    """L1 regulariser term of the reference (``OneProtLitModule.l1_regularization``)."""
    if x.numel() == 0:
        raise ValueError("mean_abs of an empty tensor")
    return _MeanAbsFn.apply(x)


class ModalitySteps:
    """``training_step`` / ``validation_step`` / ``test_step`` of ``OneProtLitModule``.

    network   ``nn.ModuleDict`` with a ``"sequence"`` encoder and one encoder per modality
              (oneprot_module.py:26); ``"seqsim"`` batches are encoded by the sequence encoder (:69-73)
    loss_fn   ``ClipLoss`` / ``SigLipLoss`` of this package (or anything with the same call)
    optimizer stepped once per modality pair, after ``clip_grad_norm_(1.0)`` (:104-106)
    metrics   optional dict ``"val_<modality>" / "test_<modality>" -> RetrievalMetric`` (:36-41)
    """

    def __init__(self, network: nn.ModuleDict, loss_fn, optimizer: Optional[torch.optim.Optimizer] = None, *,
                 use_l1_regularization: bool = False, train_on_all_modalities_after_step: int = 0, use_seqsim: bool = False,
                 metrics: Optional[Dict[str, object]] = None, gradient_clip_val: float = 1.0):
        self.network = network
        self.loss_fn = loss_fn
        self.optimizer = optimizer
        self.use_l1_regularization = use_l1_regularization
        self.train_on_all_modalities_after_step = train_on_all_modalities_after_step
        self.use_seqsim = use_seqsim
        self.metrics = metrics if metrics is not None else {}
        self.gradient_clip_val = gradient_clip_val
        self.global_step = 0            # optimizer steps taken, as Lightning counts them under manual optimisation

    def forward(self, x, modality: str = "sequence"):
        if modality in ("sequence", "seqsim"):
            modality = "sequence"
        return self.network[modality](x)

    __call__ = forward

    def l1_regularization(self, features):
        return mean_abs(features)

    def training_step(self, batch: Dict[str, tuple]):
        """-> list of the per-modality loss values (detached), in the order they were trained."""
        if self.optimizer is None:
            raise RuntimeError("ModalitySteps.training_step needs an optimizer")
        if self.global_step < self.train_on_all_modalities_after_step:
            modalities = ["struct_token"]
        else:
            modalities = list(batch.keys())
            if not self.use_seqsim and "seqsim" in modalities:
                modalities.remove("seqsim")
        losses = []
        params = [p for g in self.optimizer.param_groups for p in g["params"]]
        for modality in modalities:
            sequence_inputs, modality_inputs = batch[modality][0], batch[modality][1]
            sequence_features = self.forward(sequence_inputs, "sequence")
            modality_features = self.forward(modality_inputs, modality)
            self.optimizer.zero_grad()
            loss = self.loss_fn(sequence_features, modality_features)
            if self.use_l1_regularization:
                loss = loss + 0.01 * (mean_abs(sequence_features) + mean_abs(modality_features))
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, self.gradient_clip_val)
            self.optimizer.step()
            self.global_step += 1
            losses.append(loss.detach())
        return losses

    @torch.no_grad()
    def validation_step(self, batch):
        sequence_inputs, modality_inputs, modality = batch[0], batch[1], batch[2]
        sequence_features = self.forward(sequence_inputs, "sequence")
        modality_features = self.forward(modality_inputs, modality)
        m = self.metrics.get("val_" + modality)
        if m is not None:
            m.update(sequence_features, modality_features)
        return self.loss_fn(sequence_features, modality_features)

    @torch.no_grad()
    def test_step(self, batch: Dict[str, tuple]):
        """-> dict modality -> loss."""
        out = {}
        for modality, inputs in batch.items():
            seq_features = self.forward(inputs[0], "sequence")
            mod_features = self.forward(inputs[1], modality)
            out[modality] = self.loss_fn(seq_features, mod_features, self.network[modality].norm[1].log_logit_scale.exp())
            m = self.metrics.get("test_" + modality)
            if m is not None:
                m.update(seq_features, mod_features)
        return out
