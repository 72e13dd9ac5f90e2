#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE ITSELF (build container
only: needs /root/reference; the GPU box never runs this).

    python oracle/make_golden.py            # writes tests/golden/ + the centroid table

What runs:
* ``models.super_guessr.SuperGuessr`` imported unmodified with ``config``
  stubbed (config.py builds HF TrainingArguments at import, which needs
  ``accelerate``; only three constants are used on this path).
* ``models.proto_refiner.ProtoRefiner.forward`` unmodified, on CPU: the method
  hard-codes ``device="cuda"`` (:185,189,205,216,231,235-236), so the module's
  ``torch`` global is wrapped by a proxy whose ``tensor()`` drops the device
  argument, and the fake prototype datasets hand out tensors whose ``.to()``
  ignores the device.  ``__init__`` (S3, proto_df.csv) is bypassed with
  ``__new__``; the attributes forward reads are set by hand.
"""
from __future__ import annotations

import hashlib
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GG_REFERENCE", "/root/reference")
GOLD = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)
sys.dont_write_bytecode = True

from geoguessr_ai_b200 import synth  # noqa: E402


def import_reference():
    cfg = types.ModuleType("config")
    cfg.LABEL_SMOOTHING_CONSTANT = 65
    cfg.CLIP_EMBED_DIM = 1024
    cfg.CLIP_PRETRAINED_HEAD = "saved_models/none.model"
    cfg.__getattr__ = lambda name: "unused-" + name  # other constants: encoder side only
    sys.modules["config"] = cfg
    os.chdir(REF)
    sys.path.insert(0, REF)
    import transformers  # noqa: F401  (models.utils imports Trainer)
    import models.utils as mutils
    import models.super_guessr as sg

    # stubs for the refiner's import chain (proto_refiner.py:18-22)
    def stub(name):
        m = types.ModuleType(name)
        m.__getattr__ = lambda attr: _Any()  # e.g. boto3.client(...) at s3bucket.py:70
        m.__path__ = []
        sys.modules[name] = m
        return m

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, k):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    for name in ["timm", "timm.data", "timm.data.transforms_factory", "boto3", "botocore",
                 "botocore.config", "botocore.exceptions", "wandb", "loguru", "dotenv", "fsspec",
                 "s3fs", "accelerate"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                stub(name)
    for name, attrs in {
        "timm": ["create_model"], "timm.data": ["resolve_data_config", "resolve_model_data_config"],
        "timm.data.transforms_factory": ["create_transform"], "botocore.config": ["Config"],
        "botocore.exceptions": ["ClientError", "EndpointConnectionError", "NoCredentialsError",
                                "BotoCoreError"],
        "loguru": ["logger"], "dotenv": ["load_dotenv"],
    }.items():
        for a in attrs:
            if not hasattr(sys.modules[name], a):
                setattr(sys.modules[name], a, _Any if a[0].isupper() else _Any())
    sys.modules["botocore.exceptions"].ClientError = type("ClientError", (Exception,), {})
    import models.proto_refiner as pr
    return mutils, sg, pr


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def run_head_case(sg, model, name, B, D, seed, bf16_round, centroids):
    C = centroids.shape[0]
    emb, W, b, labels = synth.head_inputs(B, D, C, seed=seed, bf16_round=bf16_round)
    if model.cell_layer.in_features != D:
        model.cell_layer = torch.nn.Linear(D, C)
    with torch.no_grad():
        model.cell_layer.weight.copy_(W)
        model.cell_layer.bias.copy_(b)
    model.should_smooth_labels = True
    model.serving = False
    model.train()
    model.zero_grad()
    # trainer-side label derivation, main_coordinator_idun_s3.py:390-391
    import models.utils as mutils
    dist = mutils.haversine_matrix(labels, model.geocell_centroid_coords.data.t())
    labels_clf = torch.argmin(dist, dim=-1)
    out = model(embedding=emb, labels=labels, labels_clf=labels_clf)
    out.loss.backward()
    gW = model.cell_layer.weight.grad.clone()
    gb = model.cell_layer.bias.grad.clone()
    with torch.no_grad():
        logits = model.cell_layer(emb.mean(dim=1))
        top8 = torch.topk(logits, 8, dim=-1)
        s = mutils.smooth_labels(dist)
    # hard-label branch (:383)
    model.should_smooth_labels = False
    out_hard = model(embedding=emb, labels=labels, labels_clf=labels_clf)
    # serving branch (:368-369)
    model.serving = True
    model.eval()
    with torch.no_grad():
        llh_s, topk_s, emb_s = model(embedding=emb, labels_clf=labels_clf)
    assert emb_s is emb
    rows = np.linspace(0, C - 1, 48).astype(np.int64)
    np.savez_compressed(
        os.path.join(GOLD, f"head_{name}.npz"),
        B=B, D=D, seed=seed, bf16_round=int(bf16_round),
        loss=out.loss.detach().numpy(), loss_hard=out_hard.loss.detach().numpy(),
        preds_geocell=out.preds_geocell.numpy(), preds_LLH=out.preds_LLH.numpy(),
        top5_idx=out.top5_geocells.indices.numpy(), top5_val=out.top5_geocells.values.detach().numpy(),
        top8_logit_val=top8.values.numpy(), top8_logit_idx=top8.indices.numpy(),
        serving_idx=topk_s.indices.numpy(), serving_val=topk_s.values.numpy(), serving_llh=llh_s.numpy(),
        labels_clf=labels_clf.numpy(), dmin=dist.min(dim=-1)[0].numpy(), s_sum=s.sum(-1).numpy(),
        gW_rows=rows, gW_sample=gW[rows].numpy(), gW_colsum=gW.sum(0).numpy(),
        gW_abs_sum=np.float64(gW.double().abs().sum().item()), gb=gb.numpy(),
        sha_emb=sha(emb), sha_W=sha(W),
    )
    print(f"head_{name}: loss={out.loss.item():.6f} hard={out_hard.loss.item():.6f}")


class _StayTensor(torch.Tensor):
    """A tensor whose .to()/.cuda() ignore the device (the reference calls
    ``cell_emb["embedding"].to("cuda")``, proto_refiner.py:189)."""

    def to(self, *a, **k):
        return self.as_subclass(torch.Tensor)


class _FakeProtoDataset:
    """What ``self.protos[cell]`` must support (proto_refiner.py:178-199):
    ``["embedding"] -> (P_c, D)`` and ``[int] -> entry`` with count == 0 and
    centroid_lng / centroid_lat tensors (:251-252)."""

    def __init__(self, emb, xy):
        self.emb, self.xy = emb, xy

    def __getitem__(self, k):
        if isinstance(k, str):
            assert k == "embedding"
            return self.emb.as_subclass(_StayTensor)
        return {"count": torch.tensor(0), "centroid_lng": self.xy[k, 0], "centroid_lat": self.xy[k, 1],
                "indices": []}


class _TorchProxy:
    def __init__(self):
        self._t = torch

    def __getattr__(self, k):
        return getattr(self._t, k)

    def tensor(self, *a, **k):
        k.pop("device", None)
        return torch.tensor(*a, **k)


def make_reference_refiner(pr, protos, coords, topk, max_refinement=1000, temperature=1.6):
    r = pr.ProtoRefiner.__new__(pr.ProtoRefiner)
    torch.nn.Module.__init__(r)
    r.topk, r.max_refinement, r.verbose = topk, max_refinement, False
    r.temperature = torch.nn.Parameter(torch.tensor(temperature), requires_grad=False)
    r.geo_scaling = torch.nn.Parameter(torch.tensor(20.0), requires_grad=False)
    r.protos = [None if p is None else _FakeProtoDataset(p, xy) for p, xy in zip(protos, coords)]
    r.eval()
    return r


def run_refiner_case(pr, name, B, D, P, seed, centroids, topk, with_probs, jitter, missing, far_frac=0.0):
    C = centroids.shape[0]
    rng = np.random.default_rng(seed + 99)
    sizes = synth.cell_sizes(C, P, seed=seed, mode="skewed", missing_frac=missing)
    offsets, bank, xy = synth.proto_bank(sizes, D, centroids, seed=seed, jitter_deg=jitter)
    protos, coords = synth.bank_as_lists(offsets, bank, xy)
    emb = torch.from_numpy(rng.standard_normal((B, 4, D), dtype=np.float32))
    # candidates: 5 distinct cells per query, clustered around a random cell so
    # that the 1000 km guard sees both near and far refinements
    base = rng.integers(0, C, B)
    cand = np.stack([(base + rng.integers(-40, 40, B) * (j > 0)) % C for j in range(5)], 1)
    nfar = int(far_frac * B)
    if nfar:
        cand[:nfar, 1:] = rng.integers(0, C, (nfar, 4))
    cand = torch.from_numpy(cand.astype(np.int64))
    p = rng.dirichlet(np.ones(5) * 2.0, B).astype(np.float32)
    p = -np.sort(-p, axis=1)
    cprobs = torch.from_numpy(p) if with_probs else None
    initial = centroids[cand[:, 0]].clone()
    ref = make_reference_refiner(pr, protos, coords, topk)
    pr.torch = _TorchProxy()
    try:
        with torch.no_grad():
            loss, llh, cells = ref(embedding=emb, initial_preds=initial, candidate_cells=cand,
                                   candidate_probs=cprobs)
    finally:
        pr.torch = torch
    assert loss is None
    # Margins of the reference's own decisions, from the reference's own helpers on the same inputs
    # (proto_refiner.py:189-211, :216-219): per (query, candidate) the best and second-best score and the arg-best
    # prototype, the final probabilities and the guard distance -- so that a parity test can demand that every
    # disagreement sits on a tie of the reference (score gap / probability gap below the tolerance).
    q = emb.mean(dim=1)
    score = np.full((B, topk), -100000.0, np.float32)
    second = np.full((B, topk), -np.inf, np.float32)
    pidx = np.full((B, topk), -1, np.int64)
    cos_score = np.full((B, topk), -100000.0, np.float32)
    cos_second = np.full((B, topk), -np.inf, np.float32)
    cos_pidx = np.full((B, topk), -1, np.int64)
    final = np.zeros((B, topk), np.float32)
    guard_km = np.zeros((B,), np.float32)
    from preprocessing.geo_utils import haversine as ref_haversine
    with torch.no_grad():
        for i in range(B):
            for j in range(topk):
                c = int(cand[i, j])
                if protos[c] is None:
                    continue
                for arr_s, arr_2, arr_i, fn in ((score, second, pidx, lambda m, v: -ref._euclidean_distance(m, v)),
                                                (cos_score, cos_second, cos_pidx, ref._cosine_similarity)):
                    lg = fn(protos[c], q[i])
                    arr_s[i, j] = lg.max().item()
                    arr_i[i, j] = lg.argmax().item()
                    if lg.numel() > 1:
                        arr_2[i, j] = torch.topk(lg, 2).values[1].item()
            probs = ref._temperature_softmax(torch.from_numpy(score[i]))
            cp = cprobs[i, :topk] if with_probs else torch.nn.functional.one_hot(torch.tensor(0), topk).float()
            final[i] = (cp * probs).numpy()
            gidx = int(torch.argmax(torch.from_numpy(final[i])))
            c = int(cand[i, gidx])
            xy_g = torch.zeros(1, 2) if protos[c] is None else coords[c][pidx[i, gidx]].unsqueeze(0)
            guard_km[i] = ref_haversine(initial[i].unsqueeze(0), xy_g.float())[0].item()
    np.savez_compressed(
        os.path.join(GOLD, f"refiner_{name}.npz"),
        score=score, second=second, proto_idx=pidx, final_probs=final, guard_km=guard_km,
        cos_score=cos_score, cos_second=cos_second, cos_proto_idx=cos_pidx,
        B=B, D=D, P=P, seed=seed, topk=topk, with_probs=int(with_probs), jitter=jitter,
        missing=missing, far_frac=far_frac,
        preds_LLH=llh.numpy(), preds_geocell=cells.numpy(),
        cand=cand.numpy(), cprobs=(p if with_probs else np.zeros((0,), np.float32)),
        sha_bank=sha(bank), sha_emb=sha(emb),
    )
    changed = (cells != cand[:, 0]).float().mean().item()
    print(f"refiner_{name}: changed {100 * changed:.1f}% of {B}")


def main():
    os.makedirs(GOLD, exist_ok=True)
    mutils, sg, pr = import_reference()
    torch.manual_seed(0)
    model = sg.SuperGuessr(base_model=None, panorama=True, should_smooth_labels=True)
    cent = model.geocell_centroid_coords.data.clone()
    assert cent.shape == (12647, 2) and cent.dtype == torch.float32
    digest = sha(cent)
    print("centroid sha256", digest)
    assert digest.startswith("1f02b89335e5a2e7"), digest  # SURVEY 8c pin (1)
    np.save(os.path.join(REPO, "geoguessr_ai_b200", "data", "geocell_centroids.npy"), cent.numpy())

    # known-answer vectors, SURVEY 8c pins (2)-(5), re-derived from the reference here
    cities = torch.tensor([[10.75, 59.91], [2.35, 48.86], [-122.42, 37.77], [139.69, 35.69],
                           [0.0, 0.0], [179.9, -16.5]], dtype=torch.float32)
    d = mutils.haversine_matrix(cities, cent.t())
    s = mutils.smooth_labels(d)
    anti = mutils.haversine_matrix(torch.tensor([[10.0, 60.0]]), torch.tensor([[-170.0], [-60.0]]))
    r = pr.ProtoRefiner.__new__(pr.ProtoRefiner)
    torch.nn.Module.__init__(r)
    r.temperature = torch.nn.Parameter(torch.tensor(1.6), requires_grad=False)
    tsm = r._temperature_softmax(torch.tensor([-10.0, -12.0, -30.0, -1e5, -11.0]))
    m = torch.randn(7, 16, generator=torch.Generator().manual_seed(5))
    v = torch.randn(16, generator=torch.Generator().manual_seed(6))
    np.savez_compressed(
        os.path.join(GOLD, "kat.npz"),
        centroid_sha256=digest, centroid_colsum=cent.double().sum(0).numpy(),
        cities=cities.numpy(), city_idx=d.argmin(-1).numpy(), city_km=d.min(-1)[0].numpy(),
        city_s_sum=s.sum(-1).numpy(), antipode_km=anti.numpy(),
        tsm=tsm.numpy(), euclid_m=m.numpy(), euclid_v=v.numpy(),
        euclid=r._euclidean_distance(m, v).numpy(), cosine=r._cosine_similarity(m, v).numpy(),
    )
    print("city idx", d.argmin(-1).tolist(), "km", d.min(-1)[0].tolist())

    run_head_case(sg, model, "cfg1", B=64, D=1024, seed=0, bf16_round=False, centroids=cent)
    run_head_case(sg, model, "bf16_b256", B=256, D=1024, seed=1, bf16_round=True, centroids=cent)
    run_head_case(sg, model, "tinyvit_b96", B=96, D=576, seed=4, bf16_round=True, centroids=cent)
    run_head_case(sg, model, "small", B=8, D=64, seed=3, bf16_round=False, centroids=cent)

    run_refiner_case(pr, "cfg1", B=64, D=1024, P=3 * 12647, seed=0, centroids=cent, topk=5,
                     with_probs=True, jitter=0.0, missing=0.0, far_frac=0.25)
    run_refiner_case(pr, "jitter_missing", B=96, D=256, P=60000, seed=2, centroids=cent, topk=5,
                     with_probs=True, jitter=0.5, missing=0.05, far_frac=0.25)
    run_refiner_case(pr, "top3_noprobs", B=48, D=128, P=40000, seed=5, centroids=cent, topk=3,
                     with_probs=False, jitter=0.5, missing=0.02)


if __name__ == "__main__":
    main()
