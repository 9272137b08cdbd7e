"""Deterministic synthetic embeddings shared by bench.py, the tests and the
developer tools.  Neither product code nor oracle: it only draws inputs."""
from __future__ import annotations


def synthetic_pair(n: int, d: int, *, seed: int = 1234, pair_id: int = 0, rank: int = 0,
                   correlated: bool = True, temperature_into_b: bool = True,
                   dtype: str = "bf16"):
    """Per-rank synthetic embeddings: A = normalize(randn), B = normalize(A + 0.5 randn)
    (uncorrelated: B = normalize(randn)); training-faithful scaling multiplies B by 1/0.07
    (LearnableLogitScaling is applied to the modality tower, base_encoder.py:30)
    and rounds to ``dtype``.  Returns torch CPU tensors (A, B) in ``dtype``."""
    import torch
    import torch.nn.functional as F

    g = torch.Generator(device="cpu").manual_seed(seed + 1000 * pair_id + rank)
    a = F.normalize(torch.randn(n, d, generator=g, dtype=torch.float32), dim=-1)
    noise = torch.randn(n, d, generator=g, dtype=torch.float32)
    b = F.normalize(a + 0.5 * noise, dim=-1) if correlated else F.normalize(noise, dim=-1)
    if temperature_into_b:
        b = b * (1.0 / 0.07)
    td = {"bf16": torch.bfloat16, "fp32": torch.float32, "fp16": torch.float16}[dtype]
    return a.to(td), b.to(td)
