"""Oracle (test infrastructure, see oracle/__init__.py) for the SuperGuessr half
of the path.  Eager CPU/any-device PyTorch, same ATen calls as the reference.
"""
from __future__ import annotations

from collections import namedtuple

import torch
import torch.nn.functional as F
from torch import Tensor

# models/utils.py:12-17
ModelOutput = namedtuple(
    "ModelOutput", "loss loss_clf preds_LLH preds_geocell top5_geocells embedding"
)

LABEL_SMOOTHING_CONSTANT = 65  # config.py:52
EARTH_RADIUS_M = 6378137.0  # models/utils.py:55


def haversine_matrix(x: Tensor, y: Tensor) -> Tensor:
    """All-pairs great-circle distance in km.  Follows models/utils.py:39-57.

    x: (N, 2) (lng, lat) degrees; y: (2, M).  Returns (N, M) in x's dtype.
    """
    x_rad, y_rad = torch.deg2rad(x), torch.deg2rad(y)
    delta = x_rad.unsqueeze(2) - y_rad
    p = torch.cos(x_rad[:, 1]).unsqueeze(1) * torch.cos(y_rad[1, :]).unsqueeze(0)
    a = torch.sin(delta[:, 1, :] / 2) ** 2 + p * torch.sin(delta[:, 0, :] / 2) ** 2
    c = 2 * torch.arcsin(torch.sqrt(a))
    rad = torch.tensor(EARTH_RADIUS_M, dtype=c.dtype, device=c.device)
    return (rad * c) / 1000


def smooth_labels(distances: Tensor, tau: float = LABEL_SMOOTHING_CONSTANT) -> Tensor:
    """exp(-(d - rowmin d)/tau) with NaN/inf -> 0.  Follows models/utils.py:20-32."""
    adj = distances - distances.min(dim=-1, keepdim=True)[0]
    s = torch.exp(-adj / tau)
    return torch.nan_to_num(s, nan=0.0, posinf=0.0, neginf=0.0)


def soft_targets(labels: Tensor, centroids: Tensor) -> Tensor:
    """Normalised haversine soft targets, super_guessr.py:374-377."""
    d = haversine_matrix(labels, centroids.t())
    s = smooth_labels(d)
    return s / s.sum(dim=-1, keepdim=True).clamp_min(1e-12)


def fuse_headings(embedding: Tensor, panorama: bool) -> Tensor:
    """super_guessr.py:339-351 (non-hierarchical branch)."""
    return embedding.mean(dim=1) if panorama else embedding


def forward(
    embedding: Tensor,
    weight: Tensor,
    bias: Tensor,
    centroids: Tensor,
    labels: Tensor | None = None,
    labels_clf: Tensor | None = None,
    *,
    panorama: bool = True,
    should_smooth_labels: bool = True,
    serving: bool = False,
    training: bool = True,
    num_candidates: int = 5,
):
    """SuperGuessr.forward with base_model=None, super_guessr.py:336-395."""
    x = fuse_headings(embedding, panorama)
    logits = F.linear(x, weight, bias)  # :354
    probs = torch.softmax(logits, dim=-1)  # :355
    preds = torch.argmax(probs, dim=-1)  # :358
    pred_llh = torch.index_select(centroids, 0, preds)  # :359-361
    topk = torch.topk(probs, num_candidates, dim=-1)  # :365
    if (not training) and serving:  # :368-369
        return pred_llh, topk, embedding
    if should_smooth_labels and labels is not None:  # :372-380
        t = soft_targets(labels, centroids)
        log_probs = F.log_softmax(logits, dim=-1)
        loss = -(t * log_probs).sum(dim=-1).mean()
    else:  # :383
        loss = F.cross_entropy(logits, labels_clf)
    return ModelOutput(loss, loss, pred_llh, preds, topk, embedding)


def forward_backward(embedding, weight, bias, centroids, labels, labels_clf=None, **kw):
    """One train step through autograd (main_coordinator_idun_s3.py:394,423).

    Returns (ModelOutput, dW, db).
    """
    w = weight.detach().clone().requires_grad_(True)
    b = bias.detach().clone().requires_grad_(True)
    out = forward(embedding, w, b, centroids, labels, labels_clf, **kw)
    out.loss.backward()
    return out, w.grad, b.grad


def dlogits_analytic(logits: Tensor, t: Tensor) -> Tensor:
    """Gradient identity dL/dlogits = (softmax(logits) - t) / B."""
    return (torch.softmax(logits, dim=-1) - t) / logits.shape[0]


def nearest_centroid(labels: Tensor, centroids: Tensor):
    """Trainer's label derivation, main_coordinator_idun_s3.py:390-391."""
    d = haversine_matrix(labels, centroids.t())
    return torch.argmin(d, dim=-1), d
