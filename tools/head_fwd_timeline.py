#!/usr/bin/env python
"""Per-CTA timeline of gg_head_fwd at cfg2 (B=4096, D=1024, C=12647): where a launch spends its time.

    python tools/head_fwd_timeline.py [train|serve]

Stamps (ns, relative to the earliest CTA entry): 0 entry, 1 set-up done, 2 first tile's MMAs issued, 3 last MMA issued,
4 first accumulator complete, 5 last accumulator complete, 6 last tile's chunks done, 7 column groups folded,
8 flushed + ticket drawn, 9 merged a row block, 10 epilogue drained, 11 exit."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geoguessr_ai_b200 import _lib, ops, synth  # noqa: E402
from geoguessr_ai_b200.geocells import load_packaged_centroids  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
B, D, C, k = int(os.environ.get("KB_B", 4096)), 1024, 12647, 5
dev = torch.device("cuda:0")
cent = load_packaged_centroids().to(dev)
emb, W, b, labels = synth.head_inputs(B, D, C, seed=3)
x16 = ops.fuse_headings(emb.to(dev))
w16, bp = ops.prepare_head_weights(W.to(dev), b.to(dev))
for _ in range(3):
    ops.head_forward(x16, w16, bp, C, k, cent, want_logits=mode == "train")
tl = torch.zeros((148, 64), dtype=torch.int64, device=dev)
lib = _lib.load()
names = ["entry", "setup done", "first tile MMAs issued", "last MMA issued", "first acc complete", "last acc complete",
         "last chunks done", "folded", "flushed+ticket", "merged", "epilogue drained", "exit", "merge start",
         "merge first round", "merge loads done"]
for rep in range(3):
    tl.zero_()
    torch.cuda.synchronize()
    lib.gg_debug_head_fwd_timeline(tl.data_ptr())
    ops.head_forward(x16, w16, bp, C, k, cent, want_logits=mode == "train")
    torch.cuda.synchronize()
    lib.gg_debug_head_fwd_timeline(None)
    t = tl.cpu().double()
    t0 = t[:, 0][t[:, 0] > 0].min()
    print(f"--- launch {rep} ({mode}), us relative to the first CTA's entry: min / median / max over CTAs")
    for i, n in enumerate(names):
        col = t[:, i]
        col = col[col > 0]
        if col.numel() == 0:
            continue
        r = (col - t0) / 1e3
        print(f"  {i:2d} {n:26s} {r.min():8.1f} {r.median():8.1f} {r.max():8.1f}   ({col.numel()} CTAs)")

# raw per-CTA rows of the last launch: block, smid, then the stamps in us
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"timeline_{mode}.csv")
try:
    with open(out, "w") as f:
        f.write("block,smid," + ",".join(n.replace(" ", "_") for n in names) + ","
                + ",".join(f"{n}{i}" for n in ("acc", "rdy", "iss") for i in range(16)) + "\n")
        for blk in range(t.shape[0]):
            f.write(f"{blk},{int(t[blk, 15])}," + ",".join(f"{(t[blk, i] - t0) / 1e3:.2f}" if t[blk, i] > 0 else "" for i in range(len(names)))
                    + "," + ",".join(f"{(t[blk, 16 + i] - t0) / 1e3:.2f}" if t[blk, 16 + i] > 0 else "" for i in range(48)) + "\n")
except OSError:
    pass
