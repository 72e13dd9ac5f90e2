"""Work split of gg_head_bwd (csrc/head_bwd.cu, struct PairSchedule), restated in Python and checked for the
invariants the kernel relies on, over many shapes -- including the ones no GPU test runs:

  * every (tile, k-block) unit is computed exactly once;
  * a tile is either done whole by one pair, or has exactly one FINISHING part (k0 > 0, k1 == num_k: always a pair's
    last segment) and its other parts are PARKED (k1 < num_k: always a pair's first segment, so it depends on nobody);
  * the pairs the finisher waits for -- owner(first unit of the tile) .. pair - 1, filtered by holds() -- are exactly
    the pairs that park a part of that tile (a pair with an empty range is never waited for);
  * a pair parks at most once (one parking slot per pair in the workspace).
The restatement follows the C++ line by line; integer division is floor division on non-negative values in both."""
import itertools

import pytest


class PairSchedule:
    def __init__(self, pair, npairs, num_tiles, nk, all_streamk):
        self.pair, self.npairs, self.num_k = pair, npairs, nk
        self.rounds = 0 if all_streamk else num_tiles // npairs
        self.tail_base = self.rounds * npairs
        self.tail_total = (num_tiles - self.tail_base) * nk
        self.u0 = self.tail_total * pair // npairs
        self.u1 = self.tail_total * (pair + 1) // npairs

    def parks(self):
        return self.u1 > self.u0 and self.u1 % self.num_k != 0

    def tail_segments(self):
        return ((self.u1 - 1) // self.num_k - self.u0 // self.num_k + 1) if self.u1 > self.u0 else 0

    def segments(self):
        return self.rounds + self.tail_segments()

    def get(self, i):
        pk = 1 if self.parks() else 0
        if pk and i == 0:
            j = 0
        elif i - pk < self.rounds:
            return (i - pk) * self.npairs + self.pair, 0, self.num_k
        else:
            j = i - self.rounds
        t = (self.u1 - 1) // self.num_k - j
        base = t * self.num_k
        k0 = max(self.u0, base) - base
        k1 = min(self.u1, base + self.num_k) - base
        return self.tail_base + t, k0, k1

    def holds(self, q, a, b):
        q0 = self.tail_total * q // self.npairs
        q1 = self.tail_total * (q + 1) // self.npairs
        return q1 > q0 and q0 < b and q1 > a

    def owner(self, u):
        return ((u + 1) * self.npairs + self.tail_total - 1) // self.tail_total - 1


def check(num_tiles, nk, npairs, all_streamk):
    covered = {}
    parked = {}     # tile -> set of pairs that park a part of it
    finisher = {}   # tile -> (pair, contributors it waits for)
    for p in range(npairs):
        ps = PairSchedule(p, npairs, num_tiles, nk, all_streamk)
        n = ps.segments()
        parks_here = 0
        for i in range(n):
            tile, k0, k1 = ps.get(i)
            assert 0 <= tile < num_tiles and 0 <= k0 < k1 <= nk, (p, i, tile, k0, k1)
            for k in range(k0, k1):
                assert (tile, k) not in covered, f"unit {(tile, k)} computed by pairs {covered[(tile, k)]} and {p}"
                covered[(tile, k)] = p
            if k1 < nk:  # parks: must be the pair's first piece of work
                assert i == 0, f"pair {p} parks in segment {i}"
                parks_here += 1
                parked.setdefault(tile, set()).add(p)
            elif k0 > 0:  # finishes a tile other pairs started: must be the pair's last segment
                assert i == n - 1, f"pair {p} finishes tile {tile} in segment {i} of {n}"
                a = (tile - ps.tail_base) * nk
                b = a + k0
                q0 = ps.owner(a)
                assert ps.holds(q0, a, b) and q0 < p
                finisher[tile] = (p, {q for q in range(q0, p) if ps.holds(q, a, b)})
        assert parks_here <= 1
    assert len(covered) == num_tiles * nk
    for tile, (p, waits_for) in finisher.items():
        assert waits_for == parked.get(tile, set()), (tile, p, waits_for, parked.get(tile))
        assert all(q < p for q in waits_for)
    for tile, who in parked.items():
        assert tile in finisher, f"tile {tile} is parked by {who} but nobody finishes it"


@pytest.mark.parametrize("all_streamk", [False, True])
def test_pair_schedule_invariants(all_streamk):
    shapes = [(200, 64, 74), (150, 2, 74), (6, 4, 74), (50, 64, 74), (1, 64, 74), (74, 64, 74), (75, 1, 74),
              (300, 8, 74), (200, 64, 66), (13, 3, 5), (7, 7, 7), (8, 5, 3)]
    shapes += list(itertools.product((1, 2, 3, 5, 9, 33), (1, 2, 3, 16), (1, 2, 3, 8, 74)))
    for num_tiles, nk, npairs in shapes:
        check(num_tiles, nk, npairs, all_streamk)
