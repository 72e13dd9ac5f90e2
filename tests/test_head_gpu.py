"""GPU parity of the SuperGuessr path against the oracle and the golden vectors made by the reference.

Gates (BASELINE.json north star): top-k geocell indices identical except where the score (logit) gap
is below 1e-3; loss within 1e-3 relative; gradients checked against the oracle's autograd."""
import numpy as np
import pytest
import torch

from conftest import load_golden
import geoguessr_ai_b200 as gg
from geoguessr_ai_b200 import ops, synth
from oracle import super_guessr_oracle as sgo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
C = 12647


def make_model(D, cent, W, b, precision="bf16", **kw):
    m = gg.SuperGuessr(None, panorama=True, embed_dim=D, centroids=cent, precision=precision, **kw).to(DEV)
    with torch.no_grad():
        m.cell_layer.weight.copy_(W)
        m.cell_layer.bias.copy_(b)
    return m


def assert_topk_matches(idx, ref_sorted_logits, ref_sorted_idx, gap=1e-3):
    """idx (B,k) must equal the reference order except inside groups of reference logits closer than gap."""
    idx = idx.cpu().numpy()
    B, k = idx.shape
    bad = 0
    for r in range(B):
        if np.array_equal(idx[r], ref_sorted_idx[r, :k]):
            continue
        for j in range(k):
            if idx[r, j] == ref_sorted_idx[r, j]:
                continue
            # acceptable only if the returned class is one whose reference logit ties (within gap) with rank j
            pos = np.where(ref_sorted_idx[r] == idx[r, j])[0]
            if len(pos) == 0 or abs(ref_sorted_logits[r, pos[0]] - ref_sorted_logits[r, j]) >= gap:
                bad += 1
    assert bad == 0, f"{bad} top-k entries differ beyond the {gap} score-gap tolerance"


@pytest.mark.parametrize("name,precision", [("small", "bf16x3"), ("cfg1", "bf16x3"), ("bf16_b256", "bf16"),
                                            ("tinyvit_b96", "bf16"), ("bf16_b256", "bf16x3")])
def test_train_step_matches_reference_golden(name, precision, centroids):
    g = load_golden("head_" + name)
    B, D = int(g["B"]), int(g["D"])
    emb, W, b, labels = synth.head_inputs(B, D, C, seed=int(g["seed"]), bf16_round=bool(g["bf16_round"]))
    m = make_model(D, centroids, W, b, precision, should_smooth_labels=True).train()
    out = m(embedding=emb.to(DEV), labels=labels.to(DEV), labels_clf=torch.from_numpy(g["labels_clf"]).to(DEV))
    out.loss.backward()
    assert isinstance(out, gg.ModelOutput) and out.loss is out.loss_clf
    assert abs(out.loss.item() - float(g["loss"])) <= 1e-3 * float(g["loss"])  # 1e-3 relative (north star)
    assert_topk_matches(out.top5_geocells.indices, g["top8_logit_val"], g["top8_logit_idx"])
    same = out.top5_geocells.indices.cpu().numpy() == g["top5_idx"]
    np.testing.assert_allclose(out.top5_geocells.values.cpu().numpy()[same], g["top5_val"][same], rtol=2e-2, atol=1e-7)
    assert torch.equal(out.preds_geocell, out.top5_geocells.indices[:, 0])
    np.testing.assert_array_equal(out.preds_LLH.cpu().numpy(), centroids[out.preds_geocell.cpu()].numpy())
    assert out.embedding.shape == (B, 4, D)  # the un-fused input, super_guessr.py:336,393
    gW = m.cell_layer.weight.grad.cpu()
    scale = np.abs(g["gW_sample"]).max()
    np.testing.assert_allclose(gW[torch.from_numpy(g["gW_rows"])].numpy(), g["gW_sample"], atol=2e-2 * scale)
    np.testing.assert_allclose(m.cell_layer.bias.grad.cpu().numpy(), g["gb"], atol=2e-2 * np.abs(g["gb"]).max())
    # (the class-sum of dW cancels to ~1e-7 in exact arithmetic -- not a usable check under bf16 gradients;
    #  the total gradient mass is)
    assert abs(gW.double().abs().sum().item() - float(g["gW_abs_sum"])) <= 2e-2 * float(g["gW_abs_sum"])


@pytest.mark.parametrize("name,precision", [("cfg1", "bf16x3"), ("bf16_b256", "bf16")])
def test_serving_and_hard_label_paths(name, precision, centroids):
    g = load_golden("head_" + name)
    B, D = int(g["B"]), int(g["D"])
    emb, W, b, labels = synth.head_inputs(B, D, C, seed=int(g["seed"]), bf16_round=bool(g["bf16_round"]))
    m = make_model(D, centroids, W, b, precision, should_smooth_labels=False, serving=True)
    m.train()
    out = m(embedding=emb.to(DEV), labels=labels.to(DEV), labels_clf=torch.from_numpy(g["labels_clf"]).to(DEV))
    assert abs(out.loss.item() - float(g["loss_hard"])) <= 1e-3 * float(g["loss_hard"])
    m.eval()
    llh, topk, e = m(embedding=emb)  # CPU inputs are moved in eval mode (super_guessr.py:187-199); labels_clf optional
    assert_topk_matches(topk.indices, g["top8_logit_val"], g["top8_logit_idx"])
    vals, idx = topk  # tuple-unpack like torch.return_types.topk
    same = idx.cpu().numpy() == g["serving_idx"]
    np.testing.assert_allclose(vals.cpu().numpy()[same], g["serving_val"][same], rtol=2e-2, atol=1e-7)
    rows = same[:, 0]
    np.testing.assert_array_equal(llh.cpu().numpy()[rows], g["serving_llh"][rows])
    assert e.shape == emb.shape and e.is_cuda


def test_cfg2_full_size_against_oracle(centroids):
    """BASELINE configs[1]: B=4096, D=1024, bf16 operands; oracle gets the same values upcast to fp32."""
    B, D = 4096, 1024
    emb, W, b, labels = synth.head_inputs(B, D, C, seed=1, bf16_round=True)
    m = make_model(D, centroids, W, b, "bf16", should_smooth_labels=True).train()
    out = m(embedding=emb.to(DEV), labels=labels.to(DEV), labels_clf=torch.zeros(B, dtype=torch.int64, device=DEV))
    out.loss.backward()
    ref, gW, gb = sgo.forward_backward(emb, W, b, centroids, labels)
    assert abs(out.loss.item() - ref.loss.item()) <= 1e-3 * ref.loss.item()
    logits = torch.nn.functional.linear(emb.mean(1), W, b)
    top8 = torch.topk(logits, 8, -1)
    assert_topk_matches(out.top5_geocells.indices, top8.values.numpy(), top8.indices.numpy())
    err = (m.cell_layer.weight.grad.cpu() - gW).abs().max().item()
    assert err <= 2e-2 * gW.abs().max().item(), err
    errb = (m.cell_layer.bias.grad.cpu() - gb).abs().max().item()
    assert errb <= 2e-2 * gb.abs().max().item(), errb
    # size-independent properties: each gradient row sums to ~0 over classes (softmax - targets), so
    # db sums to 0 and dW's class-sum equals 0
    assert abs(m.cell_layer.bias.grad.sum().item()) < 1e-4
    # large-magnitude entries of dW element by element (the 2 %-of-max bound above says little about them)
    g_dev = m.cell_layer.weight.grad.cpu()
    big = gW.abs() >= 0.25 * gW.abs().max()
    assert int(big.sum()) > 50
    rel = ((g_dev - gW).abs() / gW.abs())[big]
    assert rel.max().item() <= 1e-2, rel.max().item()
    # determinism: same inputs -> bit-identical loss AND gradients
    g1_w, g1_b = m.cell_layer.weight.grad.clone(), m.cell_layer.bias.grad.clone()
    m.zero_grad(set_to_none=True)
    out2 = m(embedding=emb.to(DEV), labels=labels.to(DEV), labels_clf=torch.zeros(B, dtype=torch.int64, device=DEV))
    out2.loss.backward()
    assert out2.loss.item() == out.loss.item()
    assert torch.equal(m.cell_layer.weight.grad, g1_w) and torch.equal(m.cell_layer.bias.grad, g1_b)


def test_cfg4_per_gpu_shape_against_oracle(centroids):
    """BASELINE configs[3] at 8 GPUs: TinyViT D=576 head, 4096 samples per GPU (3 embedding-column tiles: the dW
    schedule's last round is cut stream-K style, partial accumulators are parked and re-added)."""
    B, D = 4096, 576
    emb, W, b, labels = synth.head_inputs(B, D, C, seed=4, bf16_round=True)
    m = make_model(D, centroids, W, b, "bf16", should_smooth_labels=True).train()
    out = m(embedding=emb.to(DEV), labels=labels.to(DEV), labels_clf=torch.zeros(B, dtype=torch.int64, device=DEV))
    out.loss.backward()
    ref, gW, gb = sgo.forward_backward(emb, W, b, centroids, labels)
    assert abs(out.loss.item() - ref.loss.item()) <= 1e-3 * ref.loss.item()
    logits = torch.nn.functional.linear(emb.mean(1), W, b)
    top8 = torch.topk(logits, 8, -1)
    assert_topk_matches(out.top5_geocells.indices, top8.values.numpy(), top8.indices.numpy())
    assert (m.cell_layer.weight.grad.cpu() - gW).abs().max().item() <= 2e-2 * gW.abs().max().item()
    assert (m.cell_layer.bias.grad.cpu() - gb).abs().max().item() <= 2e-2 * gb.abs().max().item()
    # the same gradient from the plain fp32 contraction of the kernel's own bf16 dlogits: isolates the GEMM
    # (tile schedule, parked partials) from the loss arithmetic -- must agree to fp32 accumulation noise
    x16 = ops.fuse_headings(emb.to(DEV))
    w16, bp = ops.prepare_head_weights(W.to(DEV), b.to(DEV))
    head = ops.head_forward(x16, w16, bp, C, 5, centroids.to(DEV), want_logits=True)
    dl = ops.hav_ce(head["logits"], head["lse"], labels.to(DEV), ops.centroid_unit_vectors(centroids.to(DEV)), C)[0]
    dW, _ = ops.head_backward(dl, x16, C, D, scale=1.0 / B)
    ref_dW = (dl[:, :C].float().t() @ x16.float()) / B
    assert (dW - ref_dW).abs().max().item() <= 1e-5 * ref_dW.abs().max().item() + 1e-9


def test_loss_kernel_edge_cases(centroids):
    """Exact mode (far_km=inf) == default cut-off to fp32 noise; labels on a centroid, mid-ocean, at the
    poles / date line; non-finite labels give zero targets (utils.py:31 nan_to_num semantics)."""
    B = 48
    g = torch.Generator().manual_seed(7)
    logits = (torch.randn(B, C, generator=g) * 0.4).to(torch.bfloat16)
    labels = torch.stack([torch.rand(B, generator=g) * 360 - 180, torch.rand(B, generator=g) * 140 - 60], 1)
    labels[0] = centroids[777]
    labels[1] = torch.tensor([-150.0, -55.0])
    labels[2] = torch.tensor([180.0, 0.0])
    labels[3] = torch.tensor([-180.0, 0.0])
    labels[4] = torch.tensor([0.0, 90.0])
    labels[5] = torch.tensor([0.0, -90.0])
    ldc = ops.logits_ld(C)
    lg = torch.zeros(B, ldc, dtype=torch.bfloat16)
    lg[:, :C] = logits
    lse = torch.logsumexp(logits.float(), -1)
    xyz = ops.centroid_unit_vectors(centroids.to(DEV))
    t = sgo.soft_targets(labels, centroids)
    ref_rows = -(t * torch.log_softmax(logits.float(), -1)).sum(-1)
    ref_dl = torch.softmax(logits.float(), -1) - t
    outs = {}
    for far in (float("inf"), ops.FAR_KM_DEFAULT):
        dl, rows, ncell, nkm = ops.hav_ce(lg.to(DEV), lse.to(DEV), labels.to(DEV), xyz, C, far_km=far, want_nearest=True)
        np.testing.assert_allclose(rows.cpu().numpy(), ref_rows.numpy(), rtol=1e-3)
        np.testing.assert_allclose(dl[:, :C].float().cpu().numpy(), ref_dl.numpy(), atol=4e-3)
        idx, d = sgo.nearest_centroid(labels, centroids)
        chosen = d.gather(1, ncell.cpu()[:, None])[:, 0]
        np.testing.assert_allclose(chosen.numpy(), d.min(-1)[0].numpy(), atol=2e-2)  # ties between duplicate centroids
        np.testing.assert_allclose(nkm.cpu().numpy(), d.min(-1)[0].numpy(), atol=2e-2, rtol=1e-5)
        outs[far] = (dl, rows)
    assert (outs[float("inf")][1] - outs[ops.FAR_KM_DEFAULT][1]).abs().max().item() < 1e-5
    bad = labels.clone()
    bad[0, 0] = float("nan")
    bad[1, 1] = float("inf")
    dl, rows, _, _ = ops.hav_ce(lg.to(DEV), lse.to(DEV), bad.to(DEV), xyz, C)
    assert rows[0].item() == 0.0 and rows[1].item() == 0.0  # zero targets -> zero row loss, gradient = softmax
    np.testing.assert_allclose(dl[:2, :C].float().cpu().numpy(), torch.softmax(logits[:2].float(), -1).numpy(), atol=4e-3)


@pytest.mark.parametrize("B,Cc,D,k", [(1, 8, 8, 1), (5, 200, 72, 8), (129, 257, 64, 5), (300, 1000, 576, 3)])
def test_ragged_shapes_through_c_abi(B, Cc, D, k):
    """Tile tails in every dimension (M, N, K not multiples of 128 / 256 / 64), k up to 8."""
    g = torch.Generator().manual_seed(B * 7 + Cc)
    x = (torch.randn(B, D, generator=g) * 0.5).to(torch.bfloat16)
    W = (torch.randn(Cc, D, generator=g) / D ** 0.5).to(torch.bfloat16)
    b = torch.randn(Cc, generator=g) * 0.1
    cent = torch.stack([torch.rand(Cc, generator=g) * 360 - 180, torch.rand(Cc, generator=g) * 140 - 60], 1)
    bp = torch.zeros(ops.bias_pad_len(Cc))
    bp[:Cc] = b
    out = ops.head_forward(x.to(DEV), W.to(DEV), bp.to(DEV), Cc, k, cent.to(DEV), True)
    logits = x.float() @ W.float().t() + b
    np.testing.assert_allclose(out["logits"][:, :Cc].float().cpu().numpy(), logits.numpy(), atol=3e-2)
    np.testing.assert_allclose(out["lse"].cpu().numpy(), torch.logsumexp(logits, -1).numpy(), atol=1e-4)
    tk = torch.topk(logits, min(k + 2, Cc), -1)
    assert_topk_matches(out["topk_idx"], tk.values.numpy(), tk.indices.numpy())
    dl = (torch.randn(B, Cc, generator=g) * 1e-2).to(torch.bfloat16)
    dlp = torch.full((B, ops.logits_ld(Cc)), float("nan"), dtype=torch.bfloat16)
    dlp[:, :Cc] = dl
    dW, db = ops.head_backward(dlp.to(DEV), x.to(DEV), Cc, D, scale=1.0 / B)
    ref = dl.float().t() @ x.float() / B
    np.testing.assert_allclose(dW.cpu().numpy(), ref.numpy(), atol=1e-6 + 1e-4 * ref.abs().max().item())
    np.testing.assert_allclose(db.cpu().numpy(), (dl.float().sum(0) / B).numpy(), atol=1e-6)


def test_head_backward_by_geocell_ranges_equals_one_launch():
    """The data-parallel path computes dW / db range by range (then all-reduces each): same bits as one launch."""
    from geoguessr_ai_b200 import dp_chunk_bounds

    B, Cc, D = 200, 1500, 192
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(B, D, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    dlp = torch.zeros((B, ops.logits_ld(Cc)), dtype=torch.bfloat16)
    dlp[:, :Cc] = (torch.randn(B, Cc, generator=g) * 1e-2).to(torch.bfloat16)
    dlp = dlp.to(DEV)
    dbp = torch.randn(3, ops.logits_ld(Cc), generator=g).to(DEV)  # stand-in for the loss kernel's column-sum partials
    for partials in (None, dbp):
        dW, db = ops.head_backward(dlp, x, Cc, D, scale=0.25, db_partials=partials)
        dW2 = torch.full_like(dW, float("nan"))
        db2 = torch.full_like(db, float("nan"))
        bounds = dp_chunk_bounds(Cc, 3)
        assert len(bounds) == 3
        for c0, c1 in bounds:
            ops.head_backward(dlp, x, Cc, D, scale=0.25, db_partials=partials, c_range=(c0, c1), out=(dW2, db2))
        assert torch.equal(dW, dW2) and torch.equal(db, db2)


def test_training_loop_matches_reference_optimizer_trajectory(centroids):
    """Three AdamW steps through the module (operand cache must follow the fp32 master weights)."""
    B, D = 64, 128
    emb, W, b, labels = synth.head_inputs(B, D, C, seed=5, bf16_round=False)
    m = make_model(D, centroids, W, b, "bf16x3", should_smooth_labels=True).train()
    opt = torch.optim.AdamW(m.cell_layer.parameters(), lr=1e-2)
    w = W.clone().requires_grad_(True)
    bb = b.clone().requires_grad_(True)
    ropt = torch.optim.AdamW([w, bb], lr=1e-2)
    for _ in range(3):
        opt.zero_grad()
        out = m(embedding=emb.to(DEV), labels=labels.to(DEV), labels_clf=torch.zeros(B, dtype=torch.int64, device=DEV))
        out.loss.backward()
        opt.step()
        ropt.zero_grad()
        ref = sgo.forward(emb, w, bb, centroids, labels)
        ref.loss.backward()
        ropt.step()
        assert abs(out.loss.item() - ref.loss.item()) <= 1e-3 * ref.loss.item()
    assert out.loss.item() < 9.4  # it learns


@pytest.mark.parametrize("k", [9, 12, 16, 23])
def test_num_candidates_beyond_eight(k, centroids):
    """torch.topk has no limit on k (super_guessr.py:29,365): ranks beyond 8 come from further passes of the GEMM,
    each below the previous pass's last entry -- exact on the same fp32 accumulators; serving and training paths."""
    B, D = 200, 256
    emb, W, b, labels = synth.head_inputs(B, D, C, seed=9, bf16_round=True)
    logits = torch.nn.functional.linear(emb.mean(1), W, b)
    ref = torch.topk(logits, k + 4, -1)
    probs = torch.softmax(logits, -1)
    m = make_model(D, centroids, W, b, "bf16", serving=True, num_candidates=k, should_smooth_labels=True).eval()
    llh, topk, _ = m(embedding=emb.to(DEV))
    assert topk.indices.shape == (B, k) and topk.values.shape == (B, k)
    assert_topk_matches(topk.indices, ref.values.numpy(), ref.indices.numpy())
    got = topk.values.cpu()
    want = probs.gather(1, topk.indices.cpu())
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=2e-2, atol=1e-9)
    assert bool((got[:, :-1] >= got[:, 1:]).all())  # sorted descending across the passes
    m.train()
    out = m(embedding=emb.to(DEV), labels=labels.to(DEV), labels_clf=torch.zeros(B, dtype=torch.int64, device=DEV))
    assert torch.equal(out.top5_geocells.indices, topk.indices)
    ref_out = sgo.forward(emb, W, b, centroids, labels, None)
    assert abs(out.loss.item() - ref_out.loss.item()) <= 1e-3 * ref_out.loss.item()


def test_topk_equal_to_class_count():
    """k = C on a tiny head: every class comes back, in torch.topk's order."""
    B, Cc, D = 5, 12, 8
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(B, D, generator=g)).to(torch.bfloat16)
    W = (torch.randn(Cc, D, generator=g)).to(torch.bfloat16)
    bp = torch.zeros(ops.bias_pad_len(Cc))
    cent = torch.zeros(Cc, 2)
    out = ops.head_forward(x.to(DEV), W.to(DEV), bp.to(DEV), Cc, Cc, cent.to(DEV), False)
    logits = x.float() @ W.float().t()
    ref = torch.topk(logits, Cc, -1)
    assert_topk_matches(out["topk_idx"], ref.values.numpy(), ref.indices.numpy())
    np.testing.assert_allclose(out["topk_val"].sum(-1).cpu().numpy(), np.ones(B), rtol=1e-4)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_half_precision_embeddings_are_read_directly(dtype):
    """Opt-in bf16 / fp16 embeddings (half the PCIe bytes): the fusion kernel reads them as they are and takes the
    heading mean in fp32 -- no eager torch cast."""
    B, V, D = 70, 4, 192
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(B, V, D, generator=g).to(dtype)
    x = ops.fuse_headings(emb.to(DEV))
    want = emb.float().mean(1).to(torch.bfloat16)
    diff = (x.cpu().float() - want.float()).abs()
    ulp = want.float().abs() * 2.0 ** -7 + 1e-30  # one bf16 ulp at most where the summation order rounds differently
    assert bool((diff <= ulp).all()) and (diff > 0).float().mean().item() < 0.02
    x2, sq = ops.fuse_headings(emb.to(DEV), want_sqnorm=True)
    assert torch.equal(x, x2)
    np.testing.assert_allclose(sq.cpu().numpy(), (x.cpu().float() ** 2).sum(-1).numpy(), rtol=1e-5)
    with pytest.raises(ops._lib.GeoguessrB200Error):
        ops.fuse_headings(emb.double().to(DEV))


def test_reference_nan_row_at_an_antipode_is_not_reproduced(centroids):
    """Documented difference (SURVEY 8a notes, models/utils.py:29-31,53): within ~100 m of the antipode of a
    centroid the reference's fp32 haversine term `a` can round to 1 + 1 ulp, asin gives NaN, the row minimum
    becomes NaN and nan_to_num turns the WHOLE row's targets into zeros (row loss 0).  Whether that happens
    depends on the last-ulp rounding of sin / cos (it differs between the reference on CPU and on GPU); the chord
    form used here cannot exceed its domain, so the row keeps its proper targets.  This label triggers it in the
    CPU oracle: the oracle's row loss is 0, ours is the loss of the fp64-distance targets."""
    label = torch.tensor([[-159.2899627685547, -40.63496398925781]])
    t_ref = sgo.soft_targets(label, centroids)
    if float(t_ref.sum()) != 0.0:
        pytest.skip("this host's sin/cos do not round `a` above 1 for the recorded label")
    g = torch.Generator().manual_seed(2)
    logits = (torch.randn(1, C, generator=g) * 0.4).to(torch.bfloat16)
    lg = torch.zeros(1, ops.logits_ld(C), dtype=torch.bfloat16)
    lg[:, :C] = logits
    lse = torch.logsumexp(logits.float(), -1)
    xyz = ops.centroid_unit_vectors(centroids.to(DEV))
    dl, rows, _, _ = ops.hav_ce(lg.to(DEV), lse.to(DEV), label.to(DEV), xyz, C, far_km=float("inf"))
    x, y = torch.deg2rad(label.double()), torch.deg2rad(centroids.double())
    a = torch.sin((y[:, 1] - x[:, 1]) / 2) ** 2 + torch.cos(x[:, 1]) * torch.cos(y[:, 1]) * torch.sin((y[:, 0] - x[:, 0]) / 2) ** 2
    d = 2 * 6378.137 * torch.arcsin(torch.sqrt(a.clamp(max=1.0)))
    s = torch.exp(-(d - d.min()) / 65.0)
    t64 = (s / s.sum()).float()
    want = -(t64 * torch.log_softmax(logits.float(), -1)[0]).sum()
    assert abs(rows[0].item() - want.item()) <= 1e-3 * want.item() and rows[0].item() > 1.0
