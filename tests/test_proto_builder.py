"""Prototype bank builder.

CPU: the table parsing / grouping against the reference's `ProtoDataManager` executed on a fixture
(tests/golden/proto_table.json, oracle/make_golden_protos.py); the oracle's mean against a by-hand value.
GPU: gg_build_prototypes against the oracle (fp32 means within 1e-6 relative -- the 4-term heading sum may be
associated differently -- and the bf16 bank within one bf16 ulp of the rounded oracle), then ProtoRefiner on
the built bank against the reference refiner's restatement on the oracle-built prototypes."""
import json
import os

import numpy as np
import pytest
import torch

from geoguessr_ai_b200 import proto_builder as pb
from oracle import proto_builder_oracle as pbo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proto_table.json")


def _golden():
    with open(GOLDEN) as f:
        return json.load(f)


def _columns(g):
    t = g["table"]
    return dict(geocell_index=[r["geocell_index"] for r in t],
                indices=[float("nan") if r["indices"] is None else r["indices"] for r in t],
                centroid_lng=[r["centroid_lng"] for r in t], centroid_lat=[r["centroid_lat"] for r in t])


def test_table_grouping_matches_reference_manager():
    g = _golden()
    C = g["num_cells"]
    cols = _columns(g)
    cell_off, member_off, members, coords, rows = pb.clusters_by_cell(cols["geocell_index"], cols["indices"],
                                                                      cols["centroid_lng"], cols["centroid_lat"], C)
    assert cell_off[0] == 0 and cell_off[-1] == len(rows) == len(g["table"])
    for c in range(C):
        want = g["reference_cells"][str(c)]
        lo, hi = int(cell_off[c]), int(cell_off[c + 1])
        assert hi - lo == len(want), f"cell {c}"
        for p, w in zip(range(lo, hi), want):
            assert members[member_off[p]:member_off[p + 1]].tolist() == w["indices"], (c, p)
            assert coords[p].tolist() == [np.float32(w["centroid_lng"]), np.float32(w["centroid_lat"])]
            assert g["table"][rows[p]]["cluster_id"] == w["cluster_id"]


@pytest.mark.parametrize("val,want", [("[1, 2]", [1, 2]), ("(4,5)", [4, 5]), ("7", [7]), (" 1, 2 ,x, 3 ", [1, 2, 3]),
                                      ("", []), (float("nan"), []), ([3, "4", None], [3, 4]), (9, [9])])
def test_parse_indices_forms(val, want):
    assert pb.parse_indices(val) == want


def test_oracle_mean_by_hand():
    emb = torch.arange(2 * 2 * 4, dtype=torch.float32).reshape(2, 2, 4)  # L=2, V=2, D=4
    # location 0 -> (0+4)/2.. = [2,3,4,5]; location 1 -> [10,11,12,13]; members [1, 0, 5 (out of range)] -> mean [6,7,8,9]
    assert pbo.generate_embedding(emb, [1, 0, 5]).tolist() == [6.0, 7.0, 8.0, 9.0]
    assert pbo.generate_embedding(emb, [1, 0], valid=[True, False]).tolist() == [2.0, 3.0, 4.0, 5.0]  # location 1 skipped
    assert pbo.generate_embedding(emb, [-1, 7]).tolist() == [0.0] * 4  # no valid member -> zero vector


def test_no_cpu_fallback():
    from geoguessr_ai_b200._lib import GeoguessrB200Error

    with pytest.raises(GeoguessrB200Error):
        pb.build_prototype_bank(torch.zeros(4, 4, 8), _columns(_golden()), 12, device="cpu")


def _synthetic(L, V, D, C, P, seed):
    rng = np.random.default_rng(seed)
    emb = torch.from_numpy(rng.standard_normal((L, V, D)).astype(np.float32))
    cells = rng.integers(0, C, size=P)
    lists = []
    for _ in range(P):
        n = int(rng.integers(0, 7))
        lists.append([int(i) for i in rng.integers(-1, L + 2, size=n)])  # some out of range
    valid = rng.random(L) > 0.05
    lng, lat = rng.uniform(-180, 180, P), rng.uniform(-60, 80, P)
    return emb, cells, lists, valid, lng, lat


@pytest.mark.gpu
@pytest.mark.parametrize("L,V,D,C,P", [(300, 4, 64, 40, 90), (1000, 4, 576, 200, 700), (64, 1, 1024, 10, 33)])
def test_kernel_matches_oracle(L, V, D, C, P):
    emb, cells, lists, valid, lng, lat = _synthetic(L, V, D, C, P, seed=L + P)
    table = dict(geocell_index=cells.tolist(), indices=[str(x) for x in lists], centroid_lng=lng.tolist(),
                 centroid_lat=lat.tolist())
    cell_off, bank, xy, f32, cnt = pb.build_prototype_bank(emb, table, C, valid=valid, return_f32=True)
    protos, coords = pbo.build(emb, cells, lists, lng, lat, C, valid=valid)
    want = torch.cat([p for p in protos if p is not None])
    want_xy = torch.cat(coords)
    assert cell_off.tolist() == np.concatenate([[0], np.cumsum([0 if p is None else len(p) for p in protos])]).tolist()
    assert torch.equal(xy, want_xy)
    got = f32.cpu()
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-7), (got - want).abs().max()
    b_got, b_want = bank.cpu().float(), want.bfloat16().float()
    assert ((b_got - b_want).abs() <= b_want.abs() * 2.0 ** -7 + 1e-30).all()  # at most one bf16 ulp where fp32 differs
    assert (b_got != b_want).float().mean() < 1e-3
    order = np.argsort(cells, kind="stable")
    used = [sum(1 for i in lists[r] if 0 <= i < L and valid[i]) for r in order]
    assert cnt.cpu().tolist() == used
    zero = [k for k, u in enumerate(used) if u == 0]
    assert zero and bool((got[zero] == 0).all())


@pytest.mark.gpu
def test_refiner_on_built_bank_matches_reference_restatement():
    import geoguessr_ai_b200 as gg
    from geoguessr_ai_b200.geocells import load_packaged_centroids
    from oracle import proto_refiner_oracle as pro

    cent = load_packaged_centroids()
    C = cent.shape[0]
    L, V, D, P, B = 4000, 4, 128, 3 * 1500, 96
    rng = np.random.default_rng(5)
    emb = torch.from_numpy(rng.standard_normal((L, V, D)).astype(np.float32)).bfloat16().float()
    hot = rng.choice(C, size=1500, replace=False)  # three clusters in each of 1500 cells
    cells = np.repeat(hot, 3)
    rng.shuffle(cells)
    lists = [[int(i) for i in rng.integers(0, L, size=int(rng.integers(1, 5)))] for _ in range(P)]
    lng = cent[cells, 0].numpy() + rng.uniform(-0.3, 0.3, P)
    lat = cent[cells, 1].numpy() + rng.uniform(-0.3, 0.3, P)
    table = dict(geocell_index=cells.tolist(), indices=lists, centroid_lng=lng.tolist(), centroid_lat=lat.tolist())
    dev = torch.device("cuda:0")
    cell_off, bank, xy = pb.build_prototype_bank(emb, table, C)
    refiner = gg.ProtoRefiner(topk=5, bank=(cell_off, bank, xy), report_changed=False, device=dev)

    q = torch.from_numpy(rng.standard_normal((B, V, D)).astype(np.float32)).bfloat16().float()
    cand = torch.from_numpy(np.stack([rng.choice(hot, size=5, replace=False) for _ in range(B)])).long()
    probs = torch.softmax(torch.from_numpy(rng.standard_normal((B, 5)).astype(np.float32)), -1)
    init = cent[cand[:, 0]]
    _, llh, cell = refiner(q.to(dev), init.to(dev), cand.to(dev), probs.to(dev))
    # the reference refiner on prototypes built by the oracle, rounded to the bank's storage type
    protos, coords = pbo.build(emb, cells, lists, lng, lat, C)
    protos = [None if p is None else p.bfloat16().float() for p in protos]
    _, o_llh, o_cell, _ = pro.forward(q, init, cand, probs, protos, coords, topk=5)
    # The kernel rounds the fused query to bf16, the oracle keeps the fp32 mean, and a prototype may sit one bf16 ulp
    # apart: scores move by ~2e-3 rms.  No percentage gate: every row whose refined cell differs must sit on a tie of
    # the oracle -- two prototypes of a candidate within 0.02, or the two best final probabilities within 0.01.
    score, _, second = pro.best_per_candidate(q, cand, protos, 5)
    fp = probs * torch.softmax(score.double() / 1.6, -1).float()
    srt = fp.sort(-1, descending=True).values
    tie = ((score - second) < 0.02).any(-1) | ((srt[:, 0] - srt[:, 1]) < 0.01)
    same = cell.cpu() == o_cell
    assert int((~same & ~tie).sum()) == 0, "refined cells differ without a tie in the oracle"
    assert same.float().mean().item() >= 0.9
    clean = same & ~((score - second) < 0.02).any(-1)
    assert np.allclose(llh.cpu().numpy()[clean.numpy()], o_llh.numpy()[clean.numpy()], atol=1e-5)
