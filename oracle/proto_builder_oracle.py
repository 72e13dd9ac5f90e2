"""CPU oracle of the prototype bank builder.  Test infrastructure only (tests/, smoke, bench CPU legs).

Restates models/proto_refiner.py:461-517 (`Embeddings.generate_embeddings`) with the encoder call replaced by a
lookup of stored embeddings: per member `vec = emb[idx].mean(dim=0)` (:484-486), running fp32 sum in member order
(:493-496), `sum / count` (:516); members out of range or without finite coordinates are skipped (:467-472); no
valid member -> zero vector (:499-515).  Clusters are served per geocell in table order (models/utils.py:159-181);
prototype coordinates are the row's (centroid_lng, centroid_lat) (proto_refiner.py:251-252).

Pinned: the table parsing / grouping half against `models.utils.ProtoDataManager` executed on a fixture
(oracle/make_golden_protos.py -> tests/golden/proto_table.json).  The mean itself cannot be executed from the
reference here (it needs the vision encoders): parity of that half is by restatement -- "parity unpinned" for the
fp32 summation order inside `vec.mean(dim=0)` (a 4-term sum), covered by a 1e-6 relative tolerance in the tests.
"""
import numpy as np
import torch


def generate_embedding(emb: torch.Tensor, indices, valid=None) -> torch.Tensor:
    """One prototype: emb (L,V,D) fp32, indices = the cluster's member list."""
    L = emb.shape[0]
    total, count = None, 0
    for idx in indices:
        if idx < 0 or idx >= L:
            continue
        if valid is not None and not bool(valid[idx]):
            continue
        vec = emb[idx]
        if vec.dim() == 2:
            vec = vec.mean(dim=0)
        total = vec.clone() if total is None else total.add_(vec)
        count += 1
    if count == 0:
        return torch.zeros(emb.shape[-1], dtype=torch.float32)
    return (total / count).contiguous()


def build(emb: torch.Tensor, geocell_index, member_lists, centroid_lng, centroid_lat, num_cells: int, valid=None):
    """Per geocell (ascending), clusters in table order -> (protos list[(P_c,D) | None], coords list[(P_c,2)])."""
    by_cell = {}
    for r, c in enumerate(geocell_index):
        by_cell.setdefault(int(c), []).append(r)
    protos, coords = [], []
    for c in range(num_cells):
        rows = by_cell.get(c, [])
        if not rows:
            protos.append(None)
            coords.append(torch.zeros((0, 2)))
            continue
        protos.append(torch.stack([generate_embedding(emb, member_lists[r], valid) for r in rows]))
        coords.append(torch.tensor([[np.float32(centroid_lng[r]), np.float32(centroid_lat[r])] for r in rows],
                                   dtype=torch.float32))
    return protos, coords
