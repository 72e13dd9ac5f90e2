"""CPU oracle of the reference's alternative heading fusion, `hierarchical=True`.  Test infrastructure only.

Restates models/super_guessr.py:340-345 with models/layers/positional_encoder.py:21-44 and torch's
nn.MultiheadAttention(D, 16, dropout=0.1, batch_first=True) (super_guessr.py:89-99) in plain tensor operations,
in eval mode (both dropouts are identities):

    z[b, t]   = x[b, t] + PE[b]                      PE indexes the BATCH row (positional_encoder.py:44 adds
                                                     pos_encoding[:B] of shape (B, 1, D) to (B, 4, D)): every heading of
                                                     sample b gets the same offset, and B > 1000 does not broadcast
    q, k, v   = z Wq^T + bq, z Wk^T + bk, z Wv^T + bv   (in_proj_weight = [Wq; Wk; Wv], (3D, D))
    a[b,h,t]  = softmax_t( q[b,0,h] . k[b,t,h] / sqrt(D / 16) )     only token 0 is kept (:345 output[:, 0])
    y[b]      = (sum_t a[b,h,t] v[b,t,h])_h Wo^T + bo

Pinned: oracle/make_golden_hier.py executes the reference module (SuperGuessr(hierarchical=True) in eval mode) and
commits tests/golden/hier_fusion.npz; tests/test_oracle.py-style check in tests/test_hier_oracle.py.
"""
import math

import torch

NUM_HEADS = 16  # models/super_guessr.py:14


def positional_table(max_len: int, D: int) -> torch.Tensor:
    """positional_encoder.py:21-31: (max_len, D), sin on even columns, cos on odd."""
    pe = torch.zeros(max_len, D)
    pos = torch.arange(0, max_len, dtype=torch.float).view(-1, 1)
    div = torch.exp(torch.arange(0, D, 2).float() * (-math.log(10000.0)) / D)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def fuse(x: torch.Tensor, in_proj_weight, in_proj_bias, out_proj_weight, out_proj_bias, pos_encoding=None,
         num_heads: int = NUM_HEADS) -> torch.Tensor:
    """x (B, V, D) fp32 -> (B, D): token 0 of the self-attention over the V headings."""
    B, V, D = x.shape
    if pos_encoding is None:
        pos_encoding = positional_table(1000, D)
    pe = pos_encoding.reshape(-1, D)
    if B > pe.shape[0]:
        raise RuntimeError("the reference's positional encoding indexes the batch row and holds 1000 rows")
    z = x + pe[:B].unsqueeze(1)
    Wq, Wk, Wv = in_proj_weight[:D], in_proj_weight[D:2 * D], in_proj_weight[2 * D:]
    bq, bk, bv = in_proj_bias[:D], in_proj_bias[D:2 * D], in_proj_bias[2 * D:]
    dh = D // num_heads
    q0 = (z[:, 0] @ Wq.t() + bq).view(B, num_heads, dh)
    k = (z @ Wk.t() + bk).view(B, V, num_heads, dh)
    v = (z @ Wv.t() + bv).view(B, V, num_heads, dh)
    s = torch.einsum("bhd,bthd->bht", q0, k) / math.sqrt(dh)
    a = torch.softmax(s, dim=-1)
    ctx = torch.einsum("bht,bthd->bhd", a, v).reshape(B, D)
    return ctx @ out_proj_weight.t() + out_proj_bias
