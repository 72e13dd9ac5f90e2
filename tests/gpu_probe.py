#!/usr/bin/env python
"""Bring-up probe (not a pytest): runs every kernel in isolation on cuda:0 and prints error norms
against straightforward torch math on the same (bf16-rounded) inputs.  Used with gpurun while
developing; the real parity tests are tests/test_*_gpu.py.

    timeout 600 python tests/gpu_probe.py [stage ...]
"""
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geoguessr_ai_b200 import ops, synth  # noqa: E402
from geoguessr_ai_b200.geocells import load_packaged_centroids  # noqa: E402
from oracle import proto_refiner_oracle as pro  # noqa: E402
from oracle import super_guessr_oracle as sgo  # noqa: E402

dev = torch.device("cuda:0")
STAGES = {}


def stage(fn):
    STAGES[fn.__name__] = fn
    return fn


def report(name, got, want, tol=None):
    got, want = got.float().cpu(), want.float().cpu()
    err = (got - want).abs()
    rel = err.max() / (want.abs().max() + 1e-30)
    msg = f"  {name}: max_abs={err.max().item():.3e} rel_to_max={rel.item():.3e} mean_abs={err.mean().item():.3e}"
    if tol is not None:
        msg += "  OK" if err.max().item() <= tol else f"  **FAIL** (tol {tol})"
    print(msg, flush=True)
    return err.max().item()


@stage
def fuse():
    emb = torch.randn(300, 4, 576, device=dev)
    x, sq = ops.fuse_headings(emb, want_sqnorm=True)
    ref = emb.mean(1).to(torch.bfloat16)
    report("fuse bf16", x, ref, 0)
    report("sqnorm", sq, ref.float().pow(2).sum(1), 1e-2)
    xs = ops.fuse_headings(emb, split=True)
    m = emb.mean(1)
    report("split recon", xs[:, :576].float() + xs[:, 1152:].float(), m, 1e-4)


def _head_case(B, C, D, k=5, seed=0, want_logits=True):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(B, D, generator=g) * 0.5).to(torch.bfloat16)
    W = ((torch.rand(C, D, generator=g) * 2 - 1) / D ** 0.5).to(torch.bfloat16)
    b = ((torch.rand(C, generator=g) * 2 - 1) / D ** 0.5)
    cent = torch.stack([torch.rand(C, generator=g) * 360 - 180, torch.rand(C, generator=g) * 140 - 60], 1)
    bp = torch.zeros(ops.bias_pad_len(C))
    bp[:C] = b
    out = ops.head_forward(x.to(dev), W.to(dev), bp.to(dev), C, k, cent.to(dev), want_logits)
    torch.cuda.synchronize()
    logits = x.float() @ W.float().t() + b
    return out, logits, cent


@stage
def head_fwd_small():
    for (B, C, D) in [(128, 256, 64), (200, 1000, 128), (130, 12647, 576)]:
        print(f" head_fwd B={B} C={C} D={D}")
        out, logits, cent = _head_case(B, C, D)
        report("logits", out["logits"][:, :C], logits, 2e-2)
        report("lse", out["lse"], torch.logsumexp(logits, -1), 1e-3)
        tk = torch.topk(torch.softmax(logits, -1), 5, -1)
        mism = (out["topk_idx"].cpu() != tk.indices).sum().item()
        print(f"  topk idx mismatches: {mism} / {B * 5}")
        report("topk val", out["topk_val"], tk.values, 1e-5)
        report("pred_llh", out["pred_llh"], cent[out["pred_cell"].cpu()], 0)


@stage
def head_fwd_cfg2():
    B, C, D = 4096, 12647, 1024
    t0 = time.time()
    out, logits, cent = _head_case(B, C, D)
    report("logits", out["logits"][:, :C], logits, 3e-2)
    report("lse", out["lse"], torch.logsumexp(logits, -1), 1e-3)
    tk = torch.topk(torch.softmax(logits, -1), 5, -1)
    mism = (out["topk_idx"].cpu() != tk.indices).sum().item()
    print(f"  topk idx mismatches: {mism} / {B * 5}   ({time.time() - t0:.1f}s)")
    out2, _, _ = _head_case(B, C, D, want_logits=False)
    print("  serving == training top-k:", bool((out2["topk_idx"] == out["topk_idx"]).all()))


@stage
def hav_ce():
    cent = load_packaged_centroids()
    C = cent.shape[0]
    for B, far in [(64, float("inf")), (64, ops.FAR_KM_DEFAULT), (777, ops.FAR_KM_DEFAULT)]:
        g = torch.Generator().manual_seed(B)
        logits = (torch.randn(B, C, generator=g) * 0.3).to(torch.bfloat16)
        labels = torch.stack([torch.rand(B, generator=g) * 360 - 180, torch.rand(B, generator=g) * 140 - 60], 1)
        labels[0] = cent[1234] + 1e-3
        labels[1] = torch.tensor([-150.0, -50.0])  # far from everything
        ldc = ops.logits_ld(C)
        lg = torch.zeros(B, ldc, dtype=torch.bfloat16)
        lg[:, :C] = logits
        lse = torch.logsumexp(logits.float(), -1)
        xyz = ops.centroid_unit_vectors(cent.to(dev))
        dl, loss_rows, ncell, nkm, dbp = ops.hav_ce(lg.to(dev), lse.to(dev), labels.to(dev), xyz, C, far_km=far,
                                                    want_nearest=True, want_db=True)
        torch.cuda.synchronize()
        report("db partial sum vs colsum(dlogits)", dbp.sum(0)[:C], dl[:, :C].float().sum(0), 2e-3 * B ** 0.5)
        t = sgo.soft_targets(labels, cent)
        logp = torch.log_softmax(logits.float(), -1)
        ref_rows = -(t * logp).sum(-1)
        ref_dl = torch.softmax(logits.float(), -1) - t
        print(f" hav_ce B={B} far={far:.0f}")
        report("loss_rows", loss_rows, ref_rows, 2e-3)
        report("dlogits", dl[:, :C], ref_dl, 4e-3)
        idx, d = sgo.nearest_centroid(labels, cent)
        dsel = d.gather(1, ncell.cpu().unsqueeze(1)).squeeze(1)
        report("nearest_km vs ref min", nkm, d.min(-1)[0], 2e-2)
        report("d(ref, chosen cell) - dmin", dsel, d.min(-1)[0], 2e-2)
        print("  loss mean:", ops.loss_mean(loss_rows).item(), ref_rows.mean().item())


@stage
def hard_ce():
    B, C = 100, 1003
    logits = (torch.randn(B, C) * 0.5).to(torch.bfloat16)
    y = torch.randint(0, C, (B,))
    ldc = ops.logits_ld(C)
    lg = torch.zeros(B, ldc, dtype=torch.bfloat16)
    lg[:, :C] = logits
    lse = torch.logsumexp(logits.float(), -1)
    dl, rows = ops.hard_ce(lg.to(dev), lse.to(dev), y.to(dev), C)
    ref = torch.nn.functional.cross_entropy(logits.float(), y, reduction="none")
    report("hard loss rows", rows, ref, 1e-4)
    p = torch.softmax(logits.float(), -1)
    p[torch.arange(B), y] -= 1
    report("hard dlogits", dl[:, :C], p, 4e-3)


@stage
def head_bwd():
    for (B, C, D) in [(64, 128, 256), (200, 1000, 128), (4096, 12647, 1024), (1000, 12647, 576)]:
        g = torch.Generator().manual_seed(B + C)
        dl = (torch.randn(B, C, generator=g) * 1e-2).to(torch.bfloat16)
        x = (torch.randn(B, D, generator=g) * 0.5).to(torch.bfloat16)
        ldc = ops.logits_ld(C)
        dlp = torch.full((B, ldc), float("nan"), dtype=torch.bfloat16)  # pad must never be read
        dlp[:, :C] = dl
        gs = torch.tensor(2.0, device=dev)
        dW, db = ops.head_backward(dlp.to(dev), x.to(dev), C, D, scale=1.0 / B, grad_scale=gs)
        torch.cuda.synchronize()
        ref = (dl.float().t() @ x.float()) * (2.0 / B)
        print(f" head_bwd B={B} C={C} D={D}")
        report("dW", dW, ref, 1e-5 + 1e-3 * ref.abs().max().item())
        report("db", db, dl.float().sum(0) * (2.0 / B), 1e-5)


@stage
def proto():
    cent = load_packaged_centroids()
    C = cent.shape[0]
    for (B, D, P, topk, missing) in [(64, 64, 3 * C, 5, 0.0), (300, 256, 60000, 5, 0.05), (512, 1024, 400000, 3, 0.0)]:
        sizes = synth.cell_sizes(C, P, seed=B, mode="skewed", missing_frac=missing)
        off, bank, xy = synth.proto_bank(sizes, D, cent, seed=B, dtype=torch.bfloat16, jitter_deg=0.5)
        rng = np.random.default_rng(B)
        emb = torch.from_numpy(rng.standard_normal((B, 4, D), dtype=np.float32))
        emb = emb.mean(1).to(torch.bfloat16).float().unsqueeze(1).expand(B, 4, D).contiguous()
        base = rng.integers(0, C, B)
        cand = torch.from_numpy(np.stack([(base + 3 * j) % C for j in range(5)], 1).astype(np.int64))
        cand[: B // 4, 1:] = torch.from_numpy(rng.integers(0, 64, (B // 4, 4)))  # hot cells: > 128 pairs per cell
        q16, qn = ops.fuse_headings(emb.to(dev), want_sqnorm=True)
        bank_d = bank.to(dev)
        rec = ops.proto_retrieve(q16, qn, cand.to(dev), topk, bank_d, ops.row_sqnorm_bf16(bank_d), xy.to(dev),
                                 off.to(dev), 0, C, 0)
        torch.cuda.synchronize()
        protos, coords = synth.bank_as_lists(off, bank, xy)
        score, idx, second = pro.best_per_candidate(emb, cand, protos, topk)
        rec = rec.cpu().view(B, topk, 4)
        got_idx = rec[..., 3].contiguous().view(torch.int32).view(B, topk).long()
        offl = off.long()
        ref_idx = torch.where(idx >= 0, idx + offl[cand[:, :topk]], idx)
        bad = (got_idx != ref_idx) & ((score - second) > 1e-3)
        print(f" proto B={B} D={D} P={P} topk={topk}: idx mismatches beyond 1e-3 gap: {int(bad.sum())} / {B * topk}"
              f" (raw {int((got_idx != ref_idx).sum())})")
        report("score", rec[..., 0], score, 2e-3)
        initial = cent[cand[:, 0]]
        p = torch.from_numpy(-np.sort(-rng.dirichlet(np.ones(5) * 2, B).astype(np.float32), axis=1))
        llh, cells, guess = ops.proto_refine(rec.view(-1, 4).to(dev), 1, cand.to(dev), p.to(dev), initial.to(dev),
                                             topk, 1.6, 1000.0)
        _, rl, rc, rg = pro.forward(emb, initial, cand, p, protos, coords, topk=topk)
        print(f"  refine: cell mismatches {int((cells.cpu() != rc).sum())} / {B}")
        report("refined llh", llh, rl, 1e-5)


if __name__ == "__main__":
    names = sys.argv[1:] or list(STAGES)
    print(torch.cuda.get_device_name(0), torch.version.cuda, flush=True)
    failed = 0
    for n in names:
        print(f"== {n}", flush=True)
        try:
            STAGES[n]()
            torch.cuda.synchronize()
        except Exception:
            failed += 1
            traceback.print_exc()
            print(f"== {n} RAISED", flush=True)
            try:
                torch.cuda.synchronize()
            except Exception as e:  # sticky CUDA error: nothing more can run in this process
                print("CUDA context is dead:", e)
                break
    sys.exit(1 if failed else 0)
