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


GLOBAL_BLOCK = 1024      # rows per independently seeded block of a global pair


def synthetic_global_rows(row0: int, rows: int, d: int, *, seed: int = 1234, pair_id: int = 0, correlated: bool = True,
                          temperature_into_b: bool = True, dtype: str = "bf16"):
    """Rows [row0, row0 + rows) of ONE global synthetic pair that does not depend on how it is sharded: the pair is
    drawn in blocks of GLOBAL_BLOCK rows, block k with its own generator (seed, pair_id, k), each block exactly like
    ``synthetic_pair``.  Rank r of W holds rows [r n, (r + 1) n), so the global loss of ClipLoss(local_loss=False) is
    the same number at every world size (bench.py asserts it)."""
    import torch

    k0, k1 = row0 // GLOBAL_BLOCK, -(-(row0 + rows) // GLOBAL_BLOCK)
    parts = [synthetic_pair(GLOBAL_BLOCK, d, seed=seed + 7919 * (1 + k), pair_id=pair_id, rank=0, correlated=correlated,
                            temperature_into_b=temperature_into_b, dtype=dtype) for k in range(k0, k1)]
    lo = row0 - k0 * GLOBAL_BLOCK
    return (torch.cat([p[0] for p in parts])[lo:lo + rows].contiguous(), torch.cat([p[1] for p in parts])[lo:lo + rows].contiguous())
