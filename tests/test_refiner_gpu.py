"""GPU parity of ProtoRefiner against the golden vectors (reference forward executed on CPU) and the oracle.

Gates (north star): the prototype chosen per (query, candidate) is identical unless the reference's two best scores
of that cell are closer than the score tolerance; the refined geocell is identical unless the row sits on a tie of
the reference (a candidate's prototype tie, a final-probability gap or a guard distance inside the tolerance);
refined coordinates within 1 m haversine wherever the same prototype was chosen."""
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


def assert_explained(out, ref, off, cand, topk, tol_score, tol_prob, tol_km=0.5, max_refinement=1000.0):
    """out = (llh, cells, guess, score, proto) of the CUDA path; ref = dict(score, second, proto_idx (within cell),
    final_probs, guard_km, preds_geocell, preds_LLH).  Every disagreement must sit on a tie of the reference."""
    llh, cells, guess, score, proto = (t.cpu() for t in out)
    r_score, r_second = torch.as_tensor(ref["score"]), torch.as_tensor(ref["second"])
    r_idx = torch.as_tensor(ref["proto_idx"])
    r_gidx = torch.where(r_idx >= 0, r_idx + off.long()[cand[:, :topk]], r_idx)
    np.testing.assert_allclose(score.numpy(), r_score.numpy(), atol=max(tol_score, 2e-3))
    gap = r_score - r_second
    differs = proto.long() != r_gidx
    unexplained = differs & (gap >= tol_score)
    assert int(unexplained.sum()) == 0, f"{int(unexplained.sum())} prototypes differ beyond the {tol_score} score gap"
    fp = torch.as_tensor(ref["final_probs"])
    srt = fp.sort(-1, descending=True).values
    margin = srt[:, 0] - srt[:, 1] if fp.shape[1] > 1 else torch.full((fp.shape[0],), float("inf"))
    near_guard = (torch.as_tensor(ref["guard_km"]) - max_refinement).abs() < tol_km
    row_tie = differs.any(-1) | (margin < tol_prob) | near_guard | torch.isnan(fp).any(-1)
    r_cells = torch.as_tensor(ref["preds_geocell"])
    bad = (cells != r_cells) & ~row_tie
    assert int(bad.sum()) == 0, f"{int(bad.sum())} refined geocells differ without a tie in the reference"
    # same cell through the same prototype -> same coordinates to 1 m
    same = (cells == r_cells) & ~differs.any(-1)
    if same.any():
        d = haversine_m(llh[same], torch.as_tensor(ref["preds_LLH"])[same])
        assert float(d.max()) <= 1.0, float(d.max())
    return float((cells == r_cells).float().mean()), int(differs.sum())


@pytest.mark.parametrize("name", ["cfg1", "jitter_missing", "top3_noprobs"])
@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_refiner_matches_reference_golden(name, precision, centroids, capsys):
    """The reference run is fp32.  precision="bf16x3" (hi/lo split operands) must reproduce it under the north-star
    rule itself (1e-3); rounding queries and prototypes to bf16 (2^-9 relative per element) moves a score |q - p|
    by ~1.6e-3 rms whatever D is, so in the default mode every disagreement must sit inside a 0.02 score gap /
    a 0.01 probability gap of the reference (>= 5 sigma) -- still no percentages."""
    g, off, bank, xy, emb, cand, cprobs, initial = golden_case(name, centroids)
    topk = int(g["topk"])
    r = gg.ProtoRefiner(topk=topk, bank=(off, bank, xy), device=DEV, precision=precision).eval()  # loss is 0 in train mode (:162)
    loss, llh, cells, guess, score, proto = r(emb.to(DEV), initial.to(DEV), cand.to(DEV),
                                              None if cprobs is None else cprobs.to(DEV), return_debug=True)
    assert loss is None and llh.dtype == torch.float32 and cells.dtype == torch.int64 and llh.is_cuda
    assert "Changed geocell predictions of" in capsys.readouterr().out
    tol = 1e-3 if precision == "bf16x3" else 0.02
    agree, ties = assert_explained((llh, cells, guess, score, proto), g, off, cand, topk, tol_score=tol,
                                   tol_prob=1e-3 if precision == "bf16x3" else 0.01)
    if precision == "bf16x3":
        assert ties == 0 and agree >= 0.98  # the goldens' smallest score gap is 2.9e-3: nothing to explain away


@pytest.mark.parametrize("B,D,P,topk,missing", [(64, 64, 40000, 5, 0.0), (300, 256, 60000, 5, 0.05),
                                                (512, 1024, 300000, 3, 0.02), (33, 576, 20000, 1, 0.3)])
@pytest.mark.parametrize("metric", ["l2", "cosine"])
def test_refiner_matches_oracle_on_bf16_inputs(B, D, P, topk, missing, metric, centroids):
    """Same bf16-representable bank and queries on both sides -> decisions must agree exactly (up to the
    1e-3 score-gap rule) and coordinates to 1 m; both metrics; query gather by the TMA engine and by the copy."""
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
    protos, coords = synth.bank_as_lists(off, bank, xy)
    ref_score, ref_idx, second = pro.best_per_candidate(emb, cand, protos, topk, metric=metric)
    _, o_llh, o_cell, o_guess = pro.forward(emb, initial, cand, p, protos, coords, topk=topk, metric=metric)
    # the oracle's final probabilities / guard distance for the tie rule
    T = 1.6
    fp = p[:, :topk] * torch.softmax(ref_score.double() / T, -1).float()
    gsel = fp.argmax(-1)
    xy_sel = torch.zeros(B, 2)
    for i in range(B):
        c = int(cand[i, gsel[i]])
        if protos[c] is not None:
            xy_sel[i] = coords[c][ref_idx[i, gsel[i]]]
    ref = dict(score=ref_score, second=second, proto_idx=ref_idx, final_probs=fp,
               guard_km=pro.haversine(initial, xy_sel), preds_geocell=o_cell, preds_LLH=o_llh)
    tol = 1e-3 if metric == "l2" else 2e-5  # cosine scores live in [-1, 1]: the same rule at their scale
    outs = []
    for gather4 in (True, False):
        r = gg.ProtoRefiner(topk=topk, bank=(off, bank, xy), device=DEV, report_changed=False, metric=metric)
        r.gather4 = gather4
        out = r(emb.to(DEV), initial.to(DEV), cand.to(DEV), p.to(DEV), return_debug=True)[1:]
        assert_explained(out, ref, off, cand, topk, tol_score=tol, tol_prob=1e-4 if metric == "l2" else 1e-5)
        outs.append(out)
    for a, b in zip(*outs):  # the two operand paths feed the same MMAs: bit-identical results
        assert torch.equal(a, b)


def _big_case(centroids, P, B, seed):
    Cn, D, k = centroids.shape[0], 1024, 5
    sizes = synth.cell_sizes(Cn, P, seed=0, mode="skewed")
    off = np.zeros(Cn + 1, dtype=np.int64)
    np.cumsum(sizes, out=off[1:])
    g = torch.Generator(device=DEV).manual_seed(seed)
    bank = torch.empty((P, D), dtype=torch.bfloat16, device=DEV)
    for a in range(0, P, 1 << 20):
        b = min(P, a + (1 << 20))
        bank[a:b] = torch.randn((b - a, D), device=DEV, generator=g).to(torch.bfloat16)
    cell_of = torch.repeat_interleave(torch.arange(Cn, device=DEV), torch.from_numpy(sizes).to(DEV))
    xy = centroids.to(DEV)[cell_of] + (torch.rand((P, 2), device=DEV, generator=g) - 0.5)
    off32 = torch.from_numpy(off.astype(np.int32))
    emb = torch.randn((B, D), device=DEV, generator=g).to(torch.bfloat16).float()
    cand = torch.randint(0, Cn, (B, k), device=DEV, generator=g)
    probs = torch.softmax(torch.randn((B, k), device=DEV, generator=g), -1).sort(-1, descending=True).values
    initial = centroids.to(DEV)[cand[:, 0]]
    return off, off32, bank, xy, emb, cand, probs, initial


@pytest.mark.parametrize("P,nshards", [(1_000_000, 2), (10_000_000, 8)])
def test_full_size_sharded_equals_unsharded_and_oracle_subsample(P, nshards, centroids):
    """BASELINE configs[2] / configs[4] sizes: 1 M and 10 M prototypes (10 M x 1024 bf16 = 20.5 GB: row offsets beyond
    2^32 bytes), 65 536 queries, top-5.  (a) the geocell shards of the bank on one device, merged the way the
    all-gathered records are (gg_proto_refine with nranks = shards), give bit-identical results to the unsharded
    bank; (b) a 192-query subsample agrees with the oracle under the score-gap rule."""
    from geoguessr_ai_b200 import ops, shard_cells

    Cn, B, k = centroids.shape[0], 65536, 5
    off, off32, bank, xy, emb, cand, probs, initial = _big_case(centroids, P, B, seed=3)
    full = gg.ProtoRefiner(topk=k, protos="bank", bank=(off32, bank, xy), device=DEV, report_changed=False)
    _, llh, cells, guess, score, proto = full(emb, initial, cand, probs, return_debug=True)
    recs = []
    for r in range(nshards):
        lo, hi = shard_cells(off, nshards)[r]
        sh = gg.ProtoRefiner(topk=k, protos="bank", bank=(off32, bank[off[lo]:off[hi]], xy[off[lo]:off[hi]]),
                             shard=(r, nshards), bank_is_local=True, device=DEV, report_changed=False)
        recs.append(sh.retrieve(emb, cand))
        del sh
    llh2, cells2, guess2, score2, proto2 = ops.proto_refine(torch.stack(recs), nshards, cand, probs, initial, k, 1.6,
                                                            1000.0, want_debug=True)
    assert torch.equal(proto, proto2) and torch.equal(score, score2)
    assert torch.equal(cells, cells2) and torch.equal(llh, llh2) and torch.equal(guess, guess2)
    assert int(proto.max()) > P * 0.9  # prototypes from the far end of the bank take part

    n = 192
    sel = torch.cat([torch.arange(n // 2), torch.arange(B - n // 2, B)])
    e, c = emb[sel].cpu(), cand[sel].cpu()
    protos, coords = [None] * Cn, [None] * Cn
    for cell in set(c.flatten().tolist()):
        if off[cell + 1] > off[cell]:
            protos[cell] = bank[off[cell]:off[cell + 1]].float().cpu()
            coords[cell] = xy[off[cell]:off[cell + 1]].cpu()
    ref_score, ref_idx, second = pro.best_per_candidate(e, c, protos, k)
    _, o_llh, o_cell, _ = pro.forward(e, initial[sel].cpu(), c, probs[sel].cpu(), protos, coords, topk=k)
    fp = probs[sel].cpu() * torch.softmax(ref_score.double() / 1.6, -1).float()
    gsel = fp.argmax(-1)
    xy_sel = torch.stack([coords[int(c[i, gsel[i]])][ref_idx[i, gsel[i]]] if protos[int(c[i, gsel[i]])] is not None
                          else torch.zeros(2) for i in range(n)])
    ref = dict(score=ref_score, second=second, proto_idx=ref_idx, final_probs=fp,
               guard_km=pro.haversine(initial[sel].cpu(), xy_sel), preds_geocell=o_cell, preds_LLH=o_llh)
    out = tuple(t[sel.to(DEV)] for t in (llh, cells, guess, score, proto))
    assert_explained(out, ref, torch.from_numpy(off), c, k, tol_score=1e-3, tol_prob=1e-4)


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


def test_bank_installed_on_cpu_then_moved(centroids):
    """The reference idiom is ProtoRefiner(...).to(device) (inference.py:177): a bank installed on the CPU gets its
    prototype norms when it reaches the GPU -- same results as a bank installed on the GPU directly."""
    Cn, D, B = 800, 128, 40
    sizes = synth.cell_sizes(Cn, 4000, seed=2, mode="skewed")
    off, bank, xy = synth.proto_bank(sizes, D, centroids[:Cn], seed=2, dtype=torch.bfloat16, jitter_deg=0.3)
    g = torch.Generator().manual_seed(4)
    emb = torch.randn(B, 4, D, generator=g)
    cand = torch.randint(0, Cn, (B, 5), generator=g)
    initial = centroids[cand[:, 0]].clone()
    a = gg.ProtoRefiner(topk=5, bank=(off, bank, xy), device=DEV, report_changed=False)
    b = gg.ProtoRefiner(topk=5, bank=(off, bank, xy), device="cpu", report_changed=False)
    with pytest.raises(Exception):
        b(emb, initial, cand)  # no CPU path
    b = b.to(DEV)
    ra = a(emb.to(DEV), initial.to(DEV), cand.to(DEV), return_debug=True)[1:]
    rb = b(emb.to(DEV), initial.to(DEV), cand.to(DEV), return_debug=True)[1:]
    for x, y in zip(ra, rb):
        assert torch.equal(x, y)


def test_cosine_kat_against_reference_helper():
    """The reference's own _cosine_similarity output on a 7 x 16 matrix (tests/golden/kat.npz): the kernel's
    per-cell arg-max and score with metric="cosine"."""
    k = load_golden("kat")
    m, v = torch.from_numpy(k["euclid_m"]), torch.from_numpy(k["euclid_v"])
    cos = k["cosine"]
    off = torch.tensor([0, 7], dtype=torch.int32)
    xy = torch.arange(14, dtype=torch.float32).reshape(7, 2)
    r = gg.ProtoRefiner(topk=1, bank=(off, m, xy), device=DEV, report_changed=False, metric="cosine", precision="bf16x3")
    _, llh, cells, guess, score, proto = r(v.unsqueeze(0).to(DEV), torch.zeros(1, 2, device=DEV),
                                           torch.zeros(1, 1, dtype=torch.int64, device=DEV), return_debug=True)
    assert int(proto[0, 0]) == int(cos.argmax())
    assert abs(float(score[0, 0]) - float(cos.max())) < 1e-5
