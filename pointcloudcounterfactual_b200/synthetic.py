"""Seeded synthetic point clouds shaped like the reference's data.

All inputs are generated on the CPU with ``torch.Generator().manual_seed(seed)`` so that the CPU oracle, the
reference path and the CUDA path see identical bits.  ``normalise`` follows the reference's preprocessing
(``src/data/augmentations.py:13-18``: centre, divide by the largest norm).
"""
from __future__ import annotations

import torch


def normalise(cloud: torch.Tensor) -> torch.Tensor:
    cloud = cloud - cloud.mean(dim=0, keepdim=True)
    return cloud / cloud.norm(dim=1).max()


def _gen(seed: int) -> torch.Generator:
    return torch.Generator().manual_seed(int(seed))


def s1_near(batch: int, n: int, first: int = 0) -> tuple[torch.Tensor, torch.Tensor]:
    """S1 "near" (training-like): recon = permuted reference + 2% noise.  Returns (recon, ref), each (B,n,3)."""
    recon, ref = [], []
    for b in range(first, first + batch):
        g = _gen(1000 + b)
        r = normalise(torch.randn(n, 3, generator=g) * torch.tensor([1.0, 0.6, 0.3]))
        perm = torch.randperm(n, generator=g)
        recon.append(r[perm] + 0.02 * torch.randn(n, 3, generator=g))
        ref.append(r)
    return torch.stack(recon).contiguous(), torch.stack(ref).contiguous()


def s2_far(batch: int, n: int, m: int | None = None, first: int = 0) -> tuple[torch.Tensor, torch.Tensor]:
    """S2 "far": two independent normalised Gaussian clouds, (B,n,3) and (B,m,3)."""
    m = n if m is None else m
    a, c = [], []
    for b in range(first, first + batch):
        g = _gen(3000 + b)
        a.append(normalise(torch.randn(n, 3, generator=g)))
        c.append(normalise(torch.randn(m, 3, generator=g)))
    return torch.stack(a).contiguous(), torch.stack(c).contiguous()


def s3_ties(batch: int, n: int, first: int = 0, pool: int = 1024) -> tuple[torch.Tensor, torch.Tensor]:
    """S3 "ties": coordinates on a 1/64 grid, points drawn WITH replacement from a pool (mirrors the
    reference's sampling, ``src/data/modelnet.py:43``) -- exercises the lowest-index tie rule."""
    a, c = [], []
    for b in range(first, first + batch):
        g = _gen(4000 + b)
        base = torch.round(normalise(torch.randn(pool, 3, generator=g)) * 64.0) / 64.0
        a.append(base[torch.randint(0, pool, (n,), generator=g)])
        c.append(base[torch.randint(0, pool, (n,), generator=g)])
    return torch.stack(a).contiguous(), torch.stack(c).contiguous()


def knn_xyz(batch: int, n: int, first: int = 0) -> torch.Tensor:
    """kNN xyz input: the S1 reference clouds, channels-first (B,3,n)."""
    _, ref = s1_near(batch, n, first)
    return ref.transpose(1, 2).contiguous()


def knn_features(batch: int, channels: int, n: int, seed: int = 2000) -> torch.Tensor:
    """Post-EdgeConv-like features: leaky_relu(randn(B,C,n), 0.2), channels-first."""
    g = _gen(seed)
    return torch.nn.functional.leaky_relu(torch.randn(batch, channels, n, generator=g), 0.2).contiguous()


def auction_clouds(batch: int, n: int, seed: int = 5000) -> tuple[torch.Tensor, torch.Tensor]:
    """Auction-EMD inputs in [0,1] (its documented domain, ``external/emd/README.md:17``)."""
    g = _gen(seed)
    return torch.rand(batch, n, 3, generator=g), torch.rand(batch, n, 3, generator=g)
