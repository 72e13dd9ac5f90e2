#!/usr/bin/env python
"""CPU model of the loss kernel's work balance (no GPU): how much of `hav_ce_stream_kernel`'s time is the near-row
work being unevenly spread, and what dynamic scheduling would recover.

For the bench's labels (cfg2: B = 4096, uniform lng / lat) it computes which 64-class groups hold a near cell for
every row (exactly what gg_hav_row_stats flags), prices every (row, 256-class warp slice) in issued instructions
(pass 1 + flagged groups + per-near-row overhead, from the SASS counts in profiles/r01h_sass_hot.md) and compares
  static   one warp per (96-row block, slice), as the kernel runs today: the slowest warp ends the kernel;
  dynamic  (U rows, slice) units pulled from a queue by the same 2368 warps, + 40 instructions per unit.

    python tools/hav_balance_model.py
"""
import heapq
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geoguessr_ai_b200 import synth  # noqa: E402
from geoguessr_ai_b200.geocells import load_packaged_centroids  # noqa: E402

FAR_KM, ROWS_PER_BLOCK, WARPS = 1442.0, 96, 2368  # 148 SMs x 4 CTAs x 4 warps


def unit_vectors(ll):
    lng, lat = np.deg2rad(ll[:, 0]), np.deg2rad(ll[:, 1])
    return np.stack([np.cos(lat) * np.cos(lng), np.cos(lat) * np.sin(lng), np.sin(lat)], 1)


def main():
    cent = load_packaged_centroids().numpy().astype(np.float64)
    C, B = cent.shape[0], 4096
    labels = synth.head_inputs(B, 8, 8, seed=100)[3].numpy().astype(np.float64)
    chord = np.linalg.norm(unit_vectors(labels)[:, None, :] - unit_vectors(cent)[None, :, :], axis=2)
    d = 2 * 6378.137 * np.arcsin(np.clip(chord / 2, 0, 1))
    near = d < d.min(1, keepdims=True) + FAR_KM
    cpad = (C + 255) // 256 * 256
    padded = np.zeros((B, cpad), bool)
    padded[:, :C] = near
    groups = padded.reshape(B, cpad // 64, 64).any(2)
    slices = cpad // 256
    ng = groups.reshape(B, slices, 4).sum(2)
    cost = 55 + 35 * ng + 60 * (ng > 0)
    print(f"near cells / row {near.sum(1).mean():.0f}, near groups / row {groups.sum(1).mean():.1f} of {cpad // 64}, "
          f"near (row, slice) pairs {(ng > 0).mean():.3f}")
    blocks = (B + ROWS_PER_BLOCK - 1) // ROWS_PER_BLOCK
    static = np.array([[cost[r * ROWS_PER_BLOCK:(r + 1) * ROWS_PER_BLOCK, w].sum() for w in range(slices)]
                       for r in range(blocks)]).ravel()
    print(f"static  {ROWS_PER_BLOCK}-row blocks: mean {static.mean():.0f} max {static.max():.0f} instructions per warp "
          f"(max / mean {static.max() / static.mean():.2f})")
    for rows in (8, 16, 32):
        units = [cost[r:r + rows, w].sum() + 40 for r in range(0, B, rows) for w in range(slices)]
        order = np.random.default_rng(0).permutation(len(units))
        heap = [0.0] * WARPS
        heapq.heapify(heap)
        for i in order:
            heapq.heappush(heap, heapq.heappop(heap) + units[i])
        print(f"dynamic {rows:2d}-row units: makespan {max(heap):.0f} = {max(heap) / static.max():.2f} of the static maximum")


if __name__ == "__main__":
    main()
