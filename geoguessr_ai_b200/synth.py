"""Seeded synthetic inputs for the BASELINE.json configs.

numpy PCG64 streams, so the same seed gives the same tensors in the build
container (where the golden vectors are made from the reference) and on the
GPU box.  Nothing here touches the GPU.
"""
from __future__ import annotations

import numpy as np
import torch


def head_inputs(B: int, D: int, C: int, V: int = 4, seed: int = 0, bf16_round: bool = False):
    """Embeddings (B,V,D) ~ N(0,1); nn.Linear-style W (C,D), b (C) ~ U(+-1/sqrt(D));
    labels (B,2) = (lng ~ U(-180,180), lat ~ U(-60,80)).  With ``bf16_round`` the
    embeddings' heading mean, W and b are representable in bf16 (cfg2/cfg4: the
    oracle gets identical values upcast to fp32)."""
    rng = np.random.default_rng(seed)
    emb = rng.standard_normal((B, V, D), dtype=np.float32)
    k = 1.0 / np.sqrt(D)
    W = rng.uniform(-k, k, (C, D)).astype(np.float32)
    b = rng.uniform(-k, k, (C,)).astype(np.float32)
    lng = rng.uniform(-180.0, 180.0, B).astype(np.float32)
    lat = rng.uniform(-60.0, 80.0, B).astype(np.float32)
    emb, W, b = torch.from_numpy(emb), torch.from_numpy(W), torch.from_numpy(b)
    labels = torch.from_numpy(np.stack([lng, lat], axis=1))
    if bf16_round:
        # make every heading equal to a bf16-representable row, so that the
        # fp32 mean over headings is exactly that row
        x = emb.mean(dim=1).to(torch.bfloat16).float()
        emb = x.unsqueeze(1).expand(B, V, D).contiguous()
        W = W.to(torch.bfloat16).float()
        b = b.to(torch.bfloat16).float()
    return emb, W, b, labels


def cell_sizes(C: int, P: int, seed: int = 0, mode: str = "uniform", missing_frac: float = 0.0):
    """Number of prototypes per geocell, summing to exactly P.  ``uniform``: P/C
    each (+-1); ``skewed``: lognormal weights (real cells hold 1..12 clusters,
    SURVEY 8a a10).  ``missing_frac`` of the cells get zero prototypes
    (proto_refiner.py:111-112: a cell with no proto dataset is None)."""
    rng = np.random.default_rng(seed + 1000003)
    if mode == "uniform":
        w = np.ones(C)
    else:
        w = rng.lognormal(0.0, 0.75, C)
    if missing_frac > 0:
        w[rng.random(C) < missing_frac] = 0.0
    w = w / w.sum()
    n = np.floor(w * P).astype(np.int64)
    short = P - int(n.sum())
    order = np.argsort(-(w * P - n), kind="stable")[:short]
    n[order] += 1
    if missing_frac == 0 and P >= C:
        assert n.min() >= 0
    return n


def proto_bank(sizes: np.ndarray, D: int, centroids: torch.Tensor | None = None, seed: int = 0,
               dtype=torch.float32, jitter_deg: float = 0.0):
    """CSR prototype bank sorted by cell: offsets (C+1) int32, bank (P,D) ~ N(0,1),
    coords (P,2) (lng,lat).  In the reference every cluster of a cell carries the
    cell's own centroid (geocell_manager.py:130-131); ``jitter_deg`` > 0 gives
    each prototype distinct coordinates instead, so coordinate parity is a
    real check of WHICH prototype was picked."""
    rng = np.random.default_rng(seed + 7)
    C = len(sizes)
    offsets = np.zeros(C + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    P = int(offsets[-1])
    bank = torch.from_numpy(rng.standard_normal((P, D), dtype=np.float32)).to(dtype)
    if centroids is None:
        cell_xy = np.stack([rng.uniform(-180, 180, C), rng.uniform(-60, 80, C)], 1).astype(np.float32)
    else:
        cell_xy = centroids.cpu().numpy().astype(np.float32)
    coords = np.repeat(cell_xy, sizes, axis=0)
    if jitter_deg > 0:
        coords = coords + rng.uniform(-jitter_deg, jitter_deg, coords.shape).astype(np.float32)
    return torch.from_numpy(offsets.astype(np.int32)), bank, torch.from_numpy(coords.astype(np.float32))


def bank_as_lists(offsets: torch.Tensor, bank: torch.Tensor, coords: torch.Tensor):
    """The reference's representation: a python list with one (P_c, D) tensor (or
    None) per cell, proto_refiner.py:105-112."""
    o = offsets.tolist()
    protos, xy = [], []
    for c in range(len(o) - 1):
        if o[c + 1] == o[c]:
            protos.append(None)
            xy.append(None)
        else:
            protos.append(bank[o[c]:o[c + 1]].float())
            xy.append(coords[o[c]:o[c + 1]])
    return protos, xy
