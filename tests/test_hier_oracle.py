"""`hierarchical=True` heading fusion: the oracle against the reference module executed by
oracle/make_golden_hier.py (tests/golden/hier_fusion.npz).  CPU part: pins the checker; the CUDA path
(gg_split3_bf16 / gg_linear_bf16 / gg_hier_attention) is held to the same golden in tests/test_extras_gpu.py."""
import os

import numpy as np
import pytest
import torch

from oracle import hier_fusion_oracle as hf

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hier_fusion.npz")


def test_oracle_matches_reference_module():
    g = np.load(GOLDEN)
    t = {k: torch.from_numpy(g[k]) for k in ("x", "fused", "in_proj_weight", "in_proj_bias", "out_proj_weight",
                                             "out_proj_bias", "pos_encoding")}
    got = hf.fuse(t["x"], t["in_proj_weight"], t["in_proj_bias"], t["out_proj_weight"], t["out_proj_bias"],
                  pos_encoding=t["pos_encoding"])
    assert torch.allclose(got, t["fused"], rtol=1e-5, atol=1e-6), (got - t["fused"]).abs().max()
    # the positional table the reference registers is the closed form of positional_encoder.py:21-31
    B, D = t["x"].shape[0], t["x"].shape[2]
    assert torch.allclose(hf.positional_table(1000, D)[:B], t["pos_encoding"].reshape(B, D), atol=1e-6)
    # and it indexes the batch row: every heading of a sample gets the same offset
    again = hf.fuse(t["x"], t["in_proj_weight"], t["in_proj_bias"], t["out_proj_weight"], t["out_proj_bias"])
    assert torch.allclose(again, t["fused"], rtol=1e-5, atol=1e-6)


def test_batch_beyond_the_positional_table_fails_like_the_reference():
    D = 32
    x = torch.zeros(1001, 4, D)
    w = torch.zeros(3 * D, D)
    with pytest.raises(RuntimeError):
        hf.fuse(x, w, torch.zeros(3 * D), torch.zeros(D, D), torch.zeros(D))


def test_product_module_has_the_reference_submodules():
    """Same sub-modules, hence the same state-dict keys as the reference's hierarchical model; training the branch is
    refused (eval-mode path only), and there is no CPU path."""
    import geoguessr_ai_b200 as gg

    m = gg.SuperGuessr(None, panorama=True, hierarchical=True, embed_dim=64, centroids=torch.zeros(8, 2))
    keys = set(m.state_dict().keys())
    assert {"pos_encoder.pos_encoding", "self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight",
            "self_attn.out_proj.bias", "cell_layer.weight", "cell_layer.bias", "geocell_centroid_coords"} == keys
    assert m.pos_encoder.pos_encoding.shape == (1000, 1, 64) and not m.pos_encoder.pos_encoding.requires_grad
    assert torch.allclose(m.pos_encoder.pos_encoding.reshape(1000, 64), hf.positional_table(1000, 64), atol=1e-6)
    with pytest.raises(Exception):
        m.eval()(embedding=torch.zeros(2, 4, 64))
