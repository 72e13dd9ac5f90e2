"""`hierarchical=True` heading fusion: the oracle against the reference module executed by
oracle/make_golden_hier.py (tests/golden/hier_fusion.npz).  CPU only; the CUDA kernel for this row is not built
yet (SuperGuessr(hierarchical=True) raises NotImplementedError), so this pins the checker it will be held to."""
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


def test_product_path_still_refuses_hierarchical():
    import geoguessr_ai_b200 as gg

    with pytest.raises(NotImplementedError):
        gg.SuperGuessr(None, panorama=True, hierarchical=True, embed_dim=64, centroids=torch.zeros(8, 2))
