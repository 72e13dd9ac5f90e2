"""Oracle (test infrastructure, see oracle/__init__.py) for ProtoRefiner.forward.

Restates models/proto_refiner.py:129-237 with the device made a parameter (the
reference hard-codes "cuda").  Only the executable branch of
``_within_cluster_refinement`` (``count == 0`` -> the cluster's stored
coordinates, :251-252) is restated; the other branch dereferences a
``self.dataset`` that is never assigned (:74-76, :254).

Bank representation: ``protos[c]`` is ``None`` (cell has no prototypes,
:181-187) or a ``(P_c, D)`` tensor; ``coords[c]`` is the matching ``(P_c, 2)``
tensor of (lng, lat).  Optional ``images[c][p]`` / ``image_coords[c][p]``: the
member images ``(n, D)`` / ``(n, 2)`` of prototype p of cell c.
"""
from __future__ import annotations

import torch
from torch import Tensor

EARTH_RADIUS_M = torch.tensor(6378137.0, dtype=torch.float64)  # geo_utils.py:9


def haversine(x: Tensor, y: Tensor) -> Tensor:
    """Row-wise great-circle km, preprocessing/geo_utils.py:39-54 (fp64 radius)."""
    x_rad, y_rad = torch.deg2rad(x), torch.deg2rad(y)
    delta = y_rad - x_rad
    a = (
        torch.sin(delta[:, 1] / 2) ** 2
        + torch.cos(x_rad[:, 1]) * torch.cos(y_rad[:, 1]) * torch.sin(delta[:, 0] / 2) ** 2
    )
    c = 2 * torch.arcsin(torch.sqrt(a))
    return (EARTH_RADIUS_M.to(c.device) * c) / 1000


def euclidean_distance(matrix: Tensor, vector: Tensor) -> Tensor:
    """proto_refiner.py:364-376."""
    return torch.cdist(matrix, vector.unsqueeze(0)).flatten()


def cosine_similarity(matrix: Tensor, vector: Tensor) -> Tensor:
    """proto_refiner.py:347-362 (defined by the reference, never called by it; the opt-in ``metric="cosine"``)."""
    dot_product = torch.mm(matrix, vector.unsqueeze(1))
    matrix_norm = torch.norm(matrix, dim=1).unsqueeze(1)
    vector_norm = torch.norm(vector)
    return (dot_product / (matrix_norm * vector_norm)).flatten()


def similarity(matrix: Tensor, vector: Tensor, metric: str = "l2") -> Tensor:
    """The per-cell logits of proto_refiner.py:190: ``-_euclidean_distance`` as executed ("l2"), or the cosine
    alternative the file defines."""
    return -euclidean_distance(matrix, vector) if metric == "l2" else cosine_similarity(matrix, vector)


def temperature_softmax(x: Tensor, temperature: float | Tensor) -> Tensor:
    """proto_refiner.py:378-389 -- no max-subtraction, on purpose."""
    ex = torch.exp(x / temperature)
    return ex / torch.sum(ex, axis=0)


def forward(
    embedding: Tensor,
    initial_preds: Tensor,
    candidate_cells: Tensor,
    candidate_probs: Tensor | None,
    protos: list,
    coords: list,
    *,
    topk: int = 5,
    max_refinement: float = 1000,
    temperature: float = 1.6,
    device="cpu",
    metric: str = "l2",
    images: list | None = None,
    image_coords: list | None = None,
):
    """ProtoRefiner.forward, proto_refiner.py:129-237.  Returns
    (None, preds_LLH (B,2) fp32, preds_geocell (B,) int64, guess_index (B,))."""
    assert topk <= candidate_cells.size(1)
    if embedding.dim() == 3:  # :150-151
        embedding = embedding.mean(dim=1)
    if candidate_probs is None:  # :154-156
        candidate_probs = torch.zeros_like(candidate_cells)
        candidate_probs[:, 0] = 1
    temperature = torch.tensor(float(temperature))

    guess_index, preds_llh, preds_geocell = [], [], []
    for i, (emb, candidates, c_probs) in enumerate(
        zip(embedding, candidate_cells, candidate_probs)
    ):
        top_preds, top_distances = [], []
        for cell in candidates[:topk]:
            cell_id = cell.item()
            cell_emb = protos[cell_id]
            if cell_emb is None:  # :181-187
                top_distances.append(torch.tensor(-100000, device=device))
                top_preds.append([0.0, 0.0])
                continue
            logits = similarity(cell_emb.to(device), emb, metric)  # :189-190
            top_distances.append(torch.max(logits).item())  # :193
            j = torch.argmax(logits, dim=-1).item()  # :194
            lng, lat = coords[cell_id][j, 0].item(), coords[cell_id][j, 1].item()  # :251-252
            if images is not None and images[cell_id] is not None and images[cell_id][j].shape[0] > 0:
                # _within_cluster_refinement as it is MEANT (:239-269; as written it dereferences self.dataset, which
                # is never assigned, and takes argmax of the distances = the farthest member): the cluster's member
                # image nearest to the query gives the coordinates.  Unpinned: the reference cannot execute this.
                member_logits = similarity(images[cell_id][j].to(device), emb, metric)
                jj = torch.argmax(member_logits).item()
                lng, lat = image_coords[cell_id][j][jj, 0].item(), image_coords[cell_id][j][jj, 1].item()
            top_preds.append([lng, lat])
        top_distances = torch.tensor(top_distances, device=device)  # :205
        probs = temperature_softmax(top_distances, temperature)  # :206
        final_probs = c_probs[:topk] * probs  # :210
        refined_guess = torch.argmax(final_probs).item()  # :211
        refined = torch.tensor(top_preds[refined_guess], device=device).unsqueeze(0)
        initial = initial_preds[i].unsqueeze(0)
        distance = haversine(initial, refined)[0]  # :216-219
        if distance > max_refinement:  # :220-221
            final_probs = c_probs[:topk]
        final_pred_id = torch.argmax(final_probs).item()  # :225
        guess_index.append(final_pred_id)
        preds_llh.append(top_preds[final_pred_id])
        preds_geocell.append(candidates[final_pred_id])
    guess_index = torch.tensor(guess_index, device=device)
    preds_llh = torch.tensor(preds_llh, device=device)  # fp32, :235
    preds_geocell = torch.tensor(preds_geocell, device=device)  # int64, :236
    return None, preds_llh, preds_geocell, guess_index


def best_per_candidate(embedding: Tensor, candidate_cells: Tensor, protos: list, topk: int = 5, metric: str = "l2"):
    """Stage-1 result only: per (query, candidate) the best score
    (max_p -||proto - q||, :190-193) and the arg-best prototype index within the
    cell (:194); missing cell -> (-100000, -1).  Used to check the retrieval
    kernel separately from the refinement arithmetic."""
    if embedding.dim() == 3:
        embedding = embedding.mean(dim=1)
    B = embedding.shape[0]
    score = torch.full((B, topk), -100000.0)
    idx = torch.full((B, topk), -1, dtype=torch.int64)
    second = torch.full((B, topk), -float("inf"))
    for i in range(B):
        for j in range(topk):
            c = int(candidate_cells[i, j])
            if protos[c] is None:
                continue
            logits = similarity(protos[c], embedding[i], metric)
            score[i, j] = logits.max()
            idx[i, j] = logits.argmax()
            if logits.numel() > 1:
                second[i, j] = torch.topk(logits, 2).values[1]
    return score, idx, second
