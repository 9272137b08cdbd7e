"""oneprot_b200 - B200-native (sm_100a) implementation of OneProt's ClipLoss hot path.

Public surface mirrors the reference (klemens-floege/oneprot): ``ClipLoss``, ``gather_features``, ``SigLipLoss``
(src/models/components/loss.py), ``Normalize``, ``LearnableLogitScaling``, ``BaseEncoder`` and its layers
(base_encoder.py), ``RetrievalMetric`` (retrieval_metric.py), ``ModalitySteps`` / ``mean_abs`` (the step logic and L1
term of oneprot_module.py); plus ``NormalizeAndScale`` (fused epilogue) and ``PinnedPairPrefetcher`` (H2D staging).
"""
__version__ = "0.1.0"


def __getattr__(name):
    # lazy: importing the package must not import torch-heavy modules for `build()`
    if name in ("ClipLoss", "gather_features"):
        from . import clip_loss
        return getattr(clip_loss, name)
    if name in ("Normalize", "LearnableLogitScaling", "NormalizeAndScale"):
        from . import epilogue
        return getattr(epilogue, name)
    if name == "PinnedPairPrefetcher":
        from . import prefetch
        return prefetch.PinnedPairPrefetcher
    if name in ("RetrievalMetric", "retrieval_ranks"):
        from . import retrieval
        return getattr(retrieval, name)
    if name == "SigLipLoss":
        from . import siglip_loss
        return siglip_loss.SigLipLoss
    if name in ("BaseEncoder", "LayerNorm", "Linear", "GELU", "MeanPooling", "CLSTokenPooling"):
        from . import heads
        return getattr(heads, name)
    if name in ("ModalitySteps", "mean_abs"):
        from . import module_steps
        return getattr(module_steps, name)
    raise AttributeError(name)
