#!/usr/bin/env python
"""Per-CTA timeline of gg_head_bwd at cfg2 (B=4096, D=1024, C=12647): where a launch spends its time.

    python tools/head_bwd_timeline.py

Stamps (us relative to the earliest CTA entry), per CTA: entry, set-up done, per segment (two whole tiles, then the
pair's part of the stream-K tail round) last MMA issued / accumulator complete / epilogue done, end of the wait for the
lower pair's parked partial, exit."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geoguessr_ai_b200 import _lib, ops, synth  # noqa: E402

B, D, C = int(os.environ.get("KB_B", 4096)), 1024, 12647
dev = torch.device("cuda:0")
emb, W, b, labels = synth.head_inputs(B, D, C, seed=3)
x16 = ops.fuse_headings(emb.to(dev))
ldc = ops.logits_ld(C)
dl = torch.zeros((B, ldc), dtype=torch.bfloat16, device=dev)
dl[:, :C] = (torch.randn(B, C, device=dev) * 1e-2).to(torch.bfloat16)
for _ in range(3):
    ops.head_backward(dl, x16, C, D, scale=1.0 / B)
tl = torch.zeros((148, 64), dtype=torch.int64, device=dev)
lib = _lib.load()
for rep in range(3):
    tl.zero_()
    torch.cuda.synchronize()
    lib.gg_debug_head_bwd_timeline(tl.data_ptr())
    ops.head_backward(dl, x16, C, D, scale=1.0 / B)
    torch.cuda.synchronize()
    lib.gg_debug_head_bwd_timeline(None)
    t = tl.cpu().double()
    t0 = t[:, 0][t[:, 0] > 0].min()

    def col(i, leaders=False):
        c = t[::2, i] if leaders else t[:, i]
        c = c[c > 0]
        return (c - t0) / 1e3

    def line(name, c):
        if c.numel():
            print(f"  {name:34s} {c.min():8.1f} {c.median():8.1f} {c.max():8.1f}   ({c.numel()} CTAs)")

    print(f"--- launch {rep}: us relative to the first CTA's entry: min / median / max over CTAs")
    line("entry", col(0)); line("set-up done", col(1))
    nseg = int(t[:, 4].max())
    for i in range(min(nseg, 16)):
        line(f"segment {i}: last MMA issued", col(40 + i, True))
        line(f"segment {i}: accumulator complete", col(8 + i))
        line(f"segment {i}: epilogue done", col(24 + i))
    line("parked partial of the lower pair in", col(5))
    line("last MMA issued", col(3, True)); line("exit", col(2))
    segs = t[:, 4]
    print(f"  segments per CTA: min {int(segs.min())} max {int(segs.max())}")
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "timeline_bwd.csv")
try:
    with open(out, "w") as f:
        f.write("block," + ",".join(f"s{i}" for i in range(64)) + "\n")
        for blk in range(t.shape[0]):
            f.write(f"{blk}," + ",".join(f"{(t[blk, i] - t0) / 1e3:.2f}" if (t[blk, i] > 0 and i != 4) else (str(int(t[blk, i])) if i == 4 else "") for i in range(64)) + "\n")
except OSError:
    pass
