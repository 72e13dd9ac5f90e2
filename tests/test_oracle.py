"""Pins the oracle (oracle/) against vectors produced by EXECUTING the reference
(oracle/make_golden.py, tests/golden/*.npz) and against the analytic identities
of SURVEY.md section 8c.  CPU only."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import load_golden
from geoguessr_ai_b200 import synth
from oracle import proto_refiner_oracle as pro
from oracle import super_guessr_oracle as sgo


def test_centroid_table_pin(centroids):
    kat = load_golden("kat")
    assert centroids.shape == (12647, 2) and centroids.dtype == torch.float32
    digest = hashlib.sha256(centroids.numpy().tobytes()).hexdigest()
    assert digest == str(kat["centroid_sha256"]) and digest.startswith("1f02b89335e5a2e7")
    np.testing.assert_allclose(centroids.double().sum(0).numpy(), [381654.304, 327965.138], atol=2e-3)
    assert torch.unique(centroids, dim=0).shape[0] == 6823  # duplicated classes, SURVEY 8a-note


def test_known_answers(centroids):
    kat = load_golden("kat")
    cities = torch.from_numpy(kat["cities"])
    idx, d = sgo.nearest_centroid(cities, centroids)
    assert idx.tolist() == kat["city_idx"].tolist() == [7655, 3378, 11735, 5592, 3810, 34]
    np.testing.assert_array_equal(d.min(-1)[0].numpy(), kat["city_km"])
    np.testing.assert_allclose(kat["city_km"], [0.2502, 0.3828, 63.7728, 0.0907, 582.4782, 1025.2919], atol=1e-4)
    s = sgo.smooth_labels(d)
    np.testing.assert_array_equal(s.sum(-1).numpy(), kat["city_s_sum"])
    np.testing.assert_allclose(kat["city_s_sum"], [24.5339, 33.9900, 7.6520, 75.3943, 32.9245, 11.8518], atol=1e-3)
    anti = sgo.haversine_matrix(torch.tensor([[10.0, 60.0]]), torch.tensor([[-170.0], [-60.0]]))
    np.testing.assert_array_equal(anti.numpy(), kat["antipode_km"])
    assert abs(anti.item() - 20037.5078) < 1e-2
    assert sgo.haversine_matrix(cities, cities.t()).diagonal().abs().max() == 0
    tsm = pro.temperature_softmax(torch.tensor([-10.0, -12.0, -30.0, -1e5, -11.0]), torch.tensor(1.6))
    np.testing.assert_array_equal(tsm.numpy(), kat["tsm"])
    np.testing.assert_allclose(tsm.numpy(), [0.54892, 0.15727, 2.0456e-6, 0, 0.29381], rtol=1e-4)
    e = pro.euclidean_distance(torch.from_numpy(kat["euclid_m"]), torch.from_numpy(kat["euclid_v"]))
    np.testing.assert_array_equal(e.numpy(), kat["euclid"])


@pytest.mark.parametrize("name", ["small", "cfg1", "bf16_b256", "tinyvit_b96"])
def test_head_oracle_matches_reference(name, centroids):
    g = load_golden("head_" + name)
    B, D = int(g["B"]), int(g["D"])
    emb, W, b, labels = synth.head_inputs(B, D, 12647, seed=int(g["seed"]), bf16_round=bool(g["bf16_round"]))
    assert hashlib.sha256(emb.numpy().tobytes()).hexdigest() == str(g["sha_emb"])
    assert hashlib.sha256(W.numpy().tobytes()).hexdigest() == str(g["sha_W"])
    labels_clf, dist = sgo.nearest_centroid(labels, centroids)
    np.testing.assert_array_equal(labels_clf.numpy(), g["labels_clf"])
    out, gW, gb = sgo.forward_backward(emb, W, b, centroids, labels, labels_clf)
    # same ATen calls on the same host => identical bits
    np.testing.assert_array_equal(out.loss.detach().numpy(), g["loss"])
    np.testing.assert_array_equal(out.preds_geocell.numpy(), g["preds_geocell"])
    np.testing.assert_array_equal(out.preds_LLH.numpy(), g["preds_LLH"])
    np.testing.assert_array_equal(out.top5_geocells.indices.numpy(), g["top5_idx"])
    np.testing.assert_array_equal(out.top5_geocells.values.detach().numpy(), g["top5_val"])
    np.testing.assert_array_equal(gW[torch.from_numpy(g["gW_rows"])].numpy(), g["gW_sample"])
    np.testing.assert_array_equal(gb.numpy(), g["gb"])
    hard = sgo.forward(emb, W, b, centroids, labels, labels_clf, should_smooth_labels=False)
    np.testing.assert_array_equal(hard.loss.numpy(), g["loss_hard"])
    llh, topk, e = sgo.forward(emb, W, b, centroids, None, labels_clf, serving=True, training=False)
    np.testing.assert_array_equal(topk.indices.numpy(), g["serving_idx"])
    np.testing.assert_array_equal(topk.values.numpy(), g["serving_val"])
    np.testing.assert_array_equal(llh.numpy(), g["serving_llh"])
    assert e is emb


def test_gradient_identity(centroids):
    emb, W, b, labels = synth.head_inputs(32, 64, 12647, seed=11)
    out, gW, gb = sgo.forward_backward(emb, W, b, centroids, labels)
    x = emb.mean(1)
    logits = torch.nn.functional.linear(x, W, b)
    dl = sgo.dlogits_analytic(logits, sgo.soft_targets(labels, centroids))
    np.testing.assert_allclose((dl.t() @ x).numpy(), gW.numpy(), atol=2e-8)
    np.testing.assert_allclose(dl.sum(0).numpy(), gb.numpy(), atol=2e-8)


def _refiner_case(name, centroids):
    g = load_golden("refiner_" + name)
    B, D, P, seed = int(g["B"]), int(g["D"]), int(g["P"]), int(g["seed"])
    rng = np.random.default_rng(seed + 99)
    sizes = synth.cell_sizes(12647, P, seed=seed, mode="skewed", missing_frac=float(g["missing"]))
    offsets, bank, xy = synth.proto_bank(sizes, D, centroids, seed=seed, jitter_deg=float(g["jitter"]))
    assert hashlib.sha256(bank.numpy().tobytes()).hexdigest() == str(g["sha_bank"])
    emb = torch.from_numpy(rng.standard_normal((B, 4, D), dtype=np.float32))
    assert hashlib.sha256(emb.numpy().tobytes()).hexdigest() == str(g["sha_emb"])
    cand = torch.from_numpy(g["cand"])
    cprobs = torch.from_numpy(g["cprobs"]) if int(g["with_probs"]) else None
    initial = centroids[cand[:, 0]].clone()
    return g, offsets, bank, xy, emb, cand, cprobs, initial


@pytest.mark.parametrize("name", ["cfg1", "jitter_missing", "top3_noprobs"])
def test_refiner_oracle_matches_reference(name, centroids):
    g, offsets, bank, xy, emb, cand, cprobs, initial = _refiner_case(name, centroids)
    protos, coords = synth.bank_as_lists(offsets, bank, xy)
    _, llh, cells, _ = pro.forward(emb, initial, cand, cprobs, protos, coords, topk=int(g["topk"]))
    np.testing.assert_array_equal(cells.numpy(), g["preds_geocell"])
    np.testing.assert_array_equal(llh.numpy(), g["preds_LLH"])
    assert llh.dtype == torch.float32 and cells.dtype == torch.int64


@pytest.mark.parametrize("name", ["cfg1", "jitter_missing", "top3_noprobs"])
def test_refiner_oracle_stage1_matches_reference_helpers(name, centroids):
    """Per-(query, candidate) best score / arg-best prototype of the oracle against the reference's own
    _euclidean_distance and _cosine_similarity executed on the same inputs (stored with the golden)."""
    g, offsets, bank, xy, emb, cand, cprobs, initial = _refiner_case(name, centroids)
    protos, _ = synth.bank_as_lists(offsets, bank, xy)
    k = int(g["topk"])
    for metric, ks, ki, k2 in (("l2", "score", "proto_idx", "second"), ("cosine", "cos_score", "cos_proto_idx", "cos_second")):
        score, idx, second = pro.best_per_candidate(emb, cand, protos, k, metric=metric)
        np.testing.assert_array_equal(score.numpy(), g[ks])
        np.testing.assert_array_equal(idx.numpy(), g[ki])
        np.testing.assert_array_equal(second.numpy(), g[k2])
