#!/usr/bin/env python
"""Isolated timing of the hot launchers at cfg2 (B=4096, D=1024, C=12647): N back-to-back launches between two
CUDA events.  Tuning aid (not the bench): python tools/kbench.py [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geoguessr_ai_b200 import ops, synth  # noqa: E402
from geoguessr_ai_b200.geocells import load_packaged_centroids  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B, D, C, k = int(os.environ.get("KB_B", 4096)), int(os.environ.get("KB_D", 1024)), 12647, 5
dev = torch.device("cuda:0")
cent = load_packaged_centroids().to(dev)
emb, W, b, labels = synth.head_inputs(B, D, C, seed=3)
emb, W, b, labels = emb.to(dev), W.to(dev), b.to(dev), labels.to(dev)
table = ops.centroid_unit_vectors(cent)


ONLY = os.environ.get("KB_ONLY", "")


def timed(name, fn, work=None, unit=""):
    if ONLY and ONLY not in name:
        return None
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    extra = f"  {work / ms / 1e9:.1f} {unit}" if work else ""
    print(f"{name:28s} {ms * 1e3:8.1f} us{extra}", flush=True)
    return ms


x16 = ops.fuse_headings(emb)
w16, bp = ops.prepare_head_weights(W, b)
head = ops.head_forward(x16, w16, bp, C, k, cent, want_logits=True)
stats, _, _ = ops.hav_row_stats(labels, table, C)
dl = ops.hav_ce(head["logits"], head["lse"], None, table, C, want_db=True, want_mean=True, row_stats=stats)
flops = 2.0 * B * C * D
timed("fuse_headings", lambda: ops.fuse_headings(emb), B * D * 18, "TB/s*1e3")
timed("prepare_head_weights", lambda: ops.prepare_head_weights(W, b), C * D * 6, "TB/s*1e3")
timed("fuse_and_prepare", lambda: ops.fuse_and_prepare(emb, W, b), B * D * 18 + C * D * 6, "TB/s*1e3")
timed("head_fwd train", lambda: ops.head_forward(x16, w16, bp, C, k, cent, want_logits=True), flops, "PF/s*1e3")
timed("head_fwd serve", lambda: ops.head_forward(x16, w16, bp, C, k, cent, want_logits=False), flops, "PF/s*1e3")
timed("hav_row_stats", lambda: ops.hav_row_stats(labels, table, C, out=stats))
timed("hav_ce", lambda: ops.hav_ce(head["logits"], head["lse"], None, table, C, want_db=True, want_mean=True, row_stats=stats),
      B * C * 4, "TB/s*1e3")
timed("head_bwd", lambda: ops.head_backward(dl[0], x16, C, D, 1.0 / B, db_partials=dl[4]), flops, "PF/s*1e3")
