#!/usr/bin/env python
"""Golden vector for the `hierarchical=True` heading fusion, produced by EXECUTING the reference:
`models.super_guessr.SuperGuessr(base_model=None, panorama=True, hierarchical=True, embed_dim=D)` in eval mode,
its own PositionalEncoder + nn.MultiheadAttention, then the fused query is read back through the geocell head
(the module exposes no hook for it: with the head set to [identity; 0] the first D logits ARE the fused vector).

    python oracle/make_golden_hier.py        (needs /root/reference; build container only)
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle.make_golden import GOLD, import_reference  # noqa: E402


def main():
    _, sg, _ = import_reference()
    torch.manual_seed(42)
    B, V, D = 24, 4, 64
    model = sg.SuperGuessr(base_model=None, panorama=True, hierarchical=True, embed_dim=D)
    model.eval()
    C = model.num_cells
    with torch.no_grad():  # logits[:, :D] = fused vector
        model.cell_layer.weight.zero_()
        model.cell_layer.weight[:D] = torch.eye(D)
        model.cell_layer.bias.zero_()
    x = torch.randn(B, V, D)
    got = {}

    def grab(mod, inp, out):
        got["logits"] = out.detach().clone()

    hnd = model.cell_layer.register_forward_hook(grab)
    with torch.no_grad():
        model(embedding=x, labels_clf=torch.zeros(B, dtype=torch.long))
    hnd.remove()
    fused = got["logits"][:, :D].contiguous()
    sd = model.state_dict()
    np.savez_compressed(
        os.path.join(GOLD, "hier_fusion.npz"), x=x.numpy(), fused=fused.numpy(),
        in_proj_weight=sd["self_attn.in_proj_weight"].numpy(), in_proj_bias=sd["self_attn.in_proj_bias"].numpy(),
        out_proj_weight=sd["self_attn.out_proj.weight"].numpy(), out_proj_bias=sd["self_attn.out_proj.bias"].numpy(),
        pos_encoding=sd["pos_encoder.pos_encoding"].numpy()[:B], num_cells=C)
    print("wrote hier_fusion.npz", fused.shape, "state-dict keys:", [k for k in sd if "attn" in k or "pos" in k])


if __name__ == "__main__":
    main()
