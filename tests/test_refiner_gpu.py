"""GPU parity of ProtoRefiner against the golden vectors (reference forward executed on CPU) and the oracle.

Gates: chosen prototype identical unless the two best scores of the cell are closer than 1e-3; refined
coordinates within 1 m haversine; refined geocell identical."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import load_golden
import geoguessr_ai_b200 as gg
from geoguessr_ai_b200 import synth
from oracle import proto_refiner_oracle as pro

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def haversine_m(a, b):
    return pro.haversine(a.double(), b.double()) * 1000.0


def golden_case(name, centroids):
    g = load_golden("refiner_" + name)
    B, D, P, seed = int(g["B"]), int(g["D"]), int(g["P"]), int(g["seed"])
    rng = np.random.default_rng(seed + 99)
    sizes = synth.cell_sizes(12647, P, seed=seed, mode="skewed", missing_frac=float(g["missing"]))
    off, bank, xy = synth.proto_bank(sizes, D, centroids, seed=seed, jitter_deg=float(g["jitter"]))
    assert hashlib.sha256(bank.numpy().tobytes()).hexdigest() == str(g["sha_bank"])
    emb = torch.from_numpy(rng.standard_normal((B, 4, D), dtype=np.float32))
    cand = torch.from_numpy(g["cand"])
    cprobs = torch.from_numpy(g["cprobs"]) if int(g["with_probs"]) else None
    return g, off, bank, xy, emb, cand, cprobs, centroids[cand[:, 0]].clone()


@pytest.mark.parametrize("name", ["cfg1", "jitter_missing", "top3_noprobs"])
def test_refiner_matches_reference_golden(name, centroids, capsys):
    """fp32 reference run vs. our bf16 bank/queries: decisions can only differ where the reference's own
    margins are tiny, so compare outcome-by-outcome and demand >= 97 % identical cells and, for those,
    coordinates within 1 m."""
    g, off, bank, xy, emb, cand, cprobs, initial = golden_case(name, centroids)
    r = gg.ProtoRefiner(topk=int(g["topk"]), bank=(off, bank, xy), device=DEV).eval()  # loss is 0 in train mode (:162)
    loss, llh, cells = r(emb.to(DEV), initial.to(DEV), cand.to(DEV), None if cprobs is None else cprobs.to(DEV))
    assert loss is None and llh.dtype == torch.float32 and cells.dtype == torch.int64 and llh.is_cuda
    assert "Changed geocell predictions of" in capsys.readouterr().out
    same = cells.cpu().numpy() == g["preds_geocell"]
    assert same.mean() >= 0.97, same.mean()
    d = haversine_m(llh.cpu()[same], torch.from_numpy(g["preds_LLH"])[same])
    # identical cell but a different (near-tied) prototype is possible under bf16 rounding: allow 3 %
    assert (d <= 1.0).float().mean() >= 0.97, d.max()


@pytest.mark.parametrize("B,D,P,topk,missing", [(64, 64, 40000, 5, 0.0), (300, 256, 60000, 5, 0.05),
                                                (512, 1024, 300000, 3, 0.02), (33, 576, 20000, 1, 0.3)])
def test_refiner_matches_oracle_on_bf16_inputs(B, D, P, topk, missing, centroids):
    """Same bf16-representable bank and queries on both sides -> decisions must agree exactly (up to the
    1e-3 score-gap rule) and coordinates to 1 m."""
    Cn = centroids.shape[0]
    sizes = synth.cell_sizes(Cn, P, seed=B, mode="skewed", missing_frac=missing)
    off, bank, xy = synth.proto_bank(sizes, D, centroids, seed=B, dtype=torch.bfloat16, jitter_deg=0.5)
    rng = np.random.default_rng(B)
    emb = torch.from_numpy(rng.standard_normal((B, 4, D), dtype=np.float32))
    emb = emb.mean(1).to(torch.bfloat16).float().unsqueeze(1).expand(B, 4, D).contiguous()
    base = rng.integers(0, Cn, B)
    cand = torch.from_numpy(np.stack([(base + 3 * j) % Cn for j in range(5)], 1).astype(np.int64))
    cand[: B // 4, 1:] = torch.from_numpy(rng.integers(0, 64, (B // 4, 4)))  # hot cells -> several 128-pair chunks
    p = torch.from_numpy(-np.sort(-rng.dirichlet(np.ones(5) * 2, B).astype(np.float32), axis=1))
    initial = centroids[cand[:, 0]].clone()
    r = gg.ProtoRefiner(topk=topk, bank=(off, bank, xy), device=DEV, report_changed=False)
    _, llh, cells, guess, score, proto = r(emb.to(DEV), initial.to(DEV), cand.to(DEV), p.to(DEV), return_debug=True)
    protos, coords = synth.bank_as_lists(off, bank, xy)
    ref_score, ref_idx, second = pro.best_per_candidate(emb, cand, protos, topk)
    ref_gidx = torch.where(ref_idx >= 0, ref_idx + off.long()[cand[:, :topk]], ref_idx)
    np.testing.assert_allclose(score.cpu().numpy(), ref_score.numpy(), atol=2e-3)
    mism = (proto.cpu().long() != ref_gidx) & ((ref_score - second) > 1e-3)
    assert int(mism.sum()) == 0
    _, o_llh, o_cell, o_guess = pro.forward(emb, initial, cand, p, protos, coords, topk=topk)
    agree = cells.cpu() == o_cell
    assert agree.float().mean() >= 0.995
    assert (haversine_m(llh.cpu()[agree], o_llh[agree]) <= 1.0).all()


def test_all_candidates_missing_and_default_probs(centroids):
    """Cells without prototypes score -100000 with coords (0,0) (proto_refiner.py:181-187); with every
    candidate missing the un-stabilised softmax is 0/0 = NaN and torch.argmax picks index 0 (:205-211)."""
    Cn, D, B = 500, 64, 16
    sizes = np.zeros(Cn, dtype=np.int64)
    sizes[100:200] = 4
    off, bank, xy = synth.proto_bank(sizes, D, centroids[:Cn], seed=3, dtype=torch.bfloat16)
    emb = torch.randn(B, D).to(torch.bfloat16).float()
    cand = torch.randint(0, 100, (B, 5))  # all missing
    cand[B // 2:] = torch.randint(100, 200, (B - B // 2, 5))
    initial = centroids[cand[:, 0]].clone()
    r = gg.ProtoRefiner(topk=5, bank=(off, bank, xy), device=DEV, report_changed=False)
    _, llh, cells = r(emb.to(DEV), initial.to(DEV), cand.to(DEV))  # candidate_probs=None -> one-hot on column 0
    protos, coords = synth.bank_as_lists(off, bank, xy)
    _, o_llh, o_cell, _ = pro.forward(emb, initial, cand, None, protos, coords, topk=5)
    assert torch.equal(cells.cpu(), o_cell)
    np.testing.assert_allclose(llh.cpu().numpy(), o_llh.numpy(), atol=1e-5)
    assert (llh[: B // 2] == 0).all()
