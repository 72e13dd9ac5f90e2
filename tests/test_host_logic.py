"""CPU: host-side logic -- module contract (constructor, state-dict keys, error behaviour), synthetic
input generators, geocell sharding, and the N > 1 merge protocol over gloo (world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import geoguessr_ai_b200 as gg
from geoguessr_ai_b200 import synth
from geoguessr_ai_b200.proto_refiner import shard_cells
from oracle import proto_refiner_oracle as pro


def test_superguessr_contract(centroids, capsys):
    m = gg.SuperGuessr(None, panorama=True, should_smooth_labels=True, serving=False, some_unknown_kwarg=1)
    out = capsys.readouterr().out
    assert "Not using keyword arguments: ['some_unknown_kwarg']" in out
    assert "Initialized SuperGuessr classification model with 12647 geocells." in out
    sd = m.state_dict()
    assert set(sd) == {"geocell_centroid_coords", "cell_layer.weight", "cell_layer.bias"}
    assert sd["cell_layer.weight"].shape == (12647, 1024) and sd["cell_layer.bias"].shape == (12647,)
    assert sd["geocell_centroid_coords"].shape == (12647, 2) and not m.geocell_centroid_coords.requires_grad
    assert torch.equal(m.geocell_centroid_coords.data, centroids)
    assert (m.num_cells, m.num_candidates, m.serving, m.hidden_size) == (12647, 5, False, 1024)
    with pytest.raises(AssertionError):
        m(pixel_values=None, embedding=None)
    h = gg.SuperGuessr(None, hierarchical=True, centroids=centroids[:16], embed_dim=64)  # reference sub-modules (:88-99)
    assert isinstance(h.self_attn, torch.nn.MultiheadAttention) and h.self_attn.num_heads == 16
    assert gg.ModelOutput._fields == ("loss", "loss_clf", "preds_LLH", "preds_geocell", "top5_geocells", "embedding")
    v, i = gg.TopK(1, 2)
    assert (v, i) == (1, 2)


def test_protorefiner_contract(centroids):
    sizes = synth.cell_sizes(50, 200, seed=0, mode="skewed", missing_frac=0.1)
    off, bank, xy = synth.proto_bank(sizes, 16, None, seed=0)
    r = gg.ProtoRefiner(topk=5, bank=(off, bank, xy), device="cpu")
    assert set(r.state_dict()) == {"temperature", "geo_scaling"}
    assert abs(r.temperature.item() - 1.6) < 1e-6 and r.geo_scaling.item() == 20.0
    assert r.bank.dtype == torch.bfloat16 and r.bank.shape == (200, 16)
    with pytest.raises(AssertionError):  # topk > number of candidate columns (proto_refiner.py:145)
        r(torch.zeros(2, 16), torch.zeros(2, 2), torch.zeros(2, 3, dtype=torch.int64))
    with pytest.raises(gg.ops._lib.GeoguessrB200Error):
        r(torch.zeros(2, 16), torch.zeros(2, 2), torch.zeros(2, 5, dtype=torch.int64))
    with pytest.raises(NotImplementedError):
        gg.ProtoRefiner(protos=None)
    protos, coords = synth.bank_as_lists(off, bank, xy)
    r2 = gg.ProtoRefiner(topk=3, protos=protos, coords=coords, device="cpu")
    assert torch.equal(r2.cell_off, r.cell_off) and torch.equal(r2.bank, r.bank)


def test_synth_is_deterministic():
    a = synth.head_inputs(4, 16, 10, seed=3)
    b = synth.head_inputs(4, 16, 10, seed=3)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    e, W, bb, _ = synth.head_inputs(4, 16, 10, seed=3, bf16_round=True)
    assert torch.equal(e.mean(1), e.mean(1).to(torch.bfloat16).float()) and torch.equal(W, W.to(torch.bfloat16).float())
    s = synth.cell_sizes(100, 1000, seed=1, mode="skewed", missing_frac=0.1)
    assert s.sum() == 1000 and (s == 0).any()


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_cells_partitions_every_cell_once(world):
    sizes = synth.cell_sizes(12647, 1_000_000, seed=2, mode="skewed", missing_frac=0.02)
    off = np.concatenate([[0], np.cumsum(sizes)])
    parts = shard_cells(off, world)
    assert parts[0][0] == 0 and parts[-1][1] == 12647
    assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
    loads = [off[hi] - off[lo] for lo, hi in parts]
    assert max(loads) - min(loads) <= 2 * sizes.max() + 1  # balanced by prototype count


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _merge_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    C, D, B, k = 300, 32, 40, 5
    sizes = synth.cell_sizes(C, 2000, seed=4, mode="skewed", missing_frac=0.05)
    off, bank, xy = synth.proto_bank(sizes, D, None, seed=4)
    rng = np.random.default_rng(5)
    emb = torch.from_numpy(rng.standard_normal((B, D), dtype=np.float32))
    cand = torch.from_numpy(rng.integers(0, C, (B, k)))
    protos, _ = synth.bank_as_lists(off, bank, xy)
    lo, hi = shard_cells(off.numpy().astype(np.int64), world)[rank]
    # what a rank's retrieval kernel produces: records for owned cells, -inf elsewhere
    local = [p if lo <= c < hi else None for c, p in enumerate(protos)]
    score, idx, _ = pro.best_per_candidate(emb, cand, local, k)
    owned = (cand >= lo) & (cand < hi)
    score = torch.where(owned, score, torch.full_like(score, -float("inf")))
    rec = torch.stack([score, idx.float()], -1)
    gathered = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(gathered, rec)
    g = torch.stack(gathered)  # (world, B, k, 2): the layout gg_proto_refine consumes
    best = g[..., 0].argmax(0)
    merged = torch.gather(g, 0, best[None, ..., None].expand(1, B, k, 2))[0]
    if rank == 0:
        full_score, full_idx, _ = pro.best_per_candidate(emb, cand, protos, k)
        q.put((torch.equal(merged[..., 0], full_score), torch.equal(merged[..., 1].long(), full_idx),
               int((g[..., 0] > -float("inf")).sum(0).max())))
    dist.destroy_process_group()


def test_sharded_retrieval_merge_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_merge_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same_score, same_idx, owners = q.get(timeout=120)
    for p in procs:
        p.join(60)
    assert same_score and same_idx and owners == 1  # exactly one rank owns each pair's cell


def _split_worker(rank, world, port, q):
    """split_queries=True exchange (ProtoRefiner.forward): all-gather queries, retrieve on the shard for all of
    them, all-to-all the records back to the query owners, merge by selection."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    C, D, Bl, k = 300, 32, 24, 5
    sizes = synth.cell_sizes(C, 2000, seed=4, mode="skewed", missing_frac=0.05)
    off, bank, xy = synth.proto_bank(sizes, D, None, seed=4)
    rng = np.random.default_rng(5)
    emb_all = torch.from_numpy(rng.standard_normal((world * Bl, D), dtype=np.float32))
    cand_all_ref = torch.from_numpy(rng.integers(0, C, (world * Bl, k)))
    emb, cand = emb_all[rank * Bl:(rank + 1) * Bl].contiguous(), cand_all_ref[rank * Bl:(rank + 1) * Bl].contiguous()
    protos, _ = synth.bank_as_lists(off, bank, xy)
    lo, hi = shard_cells(off.numpy().astype(np.int64), world)[rank]
    local = [p if lo <= c < hi else None for c, p in enumerate(protos)]
    # forward(): gather queries + candidates
    q_all, c_all = torch.empty(world * Bl, D), torch.empty(world * Bl, k, dtype=torch.int64)
    dist.all_gather_into_tensor(q_all, emb)
    dist.all_gather_into_tensor(c_all, cand)
    score, idx, _ = pro.best_per_candidate(q_all, c_all, local, k)  # what the retrieval kernel leaves on this shard
    owned = (c_all >= lo) & (c_all < hi)
    score = torch.where(owned, score, torch.full_like(score, -float("inf")))
    rec_all = torch.stack([score, idx.float()], -1).reshape(world * Bl * k, 2).contiguous()
    rec = torch.empty_like(rec_all)
    dist.all_to_all_single(rec, rec_all)
    g = rec.view(world, Bl, k, 2)  # the layout gg_proto_refine consumes for this rank's queries
    best = g[..., 0].argmax(0)
    merged = torch.gather(g, 0, best[None, ..., None].expand(1, Bl, k, 2))[0]
    full_score, full_idx, _ = pro.best_per_candidate(emb, cand, protos, k)
    ok = torch.equal(q_all, emb_all) and torch.equal(merged[..., 0], full_score) and torch.equal(merged[..., 1].long(), full_idx)
    res = torch.tensor([1 if ok else 0])
    dist.all_reduce(res)
    if rank == 0:
        q.put(int(res))
    dist.destroy_process_group()


def test_split_query_exchange_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_split_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    oks = q.get(timeout=120)
    for p in procs:
        p.join(60)
    assert oks == 2  # both ranks recover, for their own queries, exactly the unsharded result


def _dp_worker(rank, world, port, q):
    """Data-parallel head step: averaged per-rank gradients == gradient of the global batch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import super_guessr_oracle as sgo

    C, D, B = 200, 16, 12
    cent = torch.stack([torch.linspace(-170, 170, C), torch.linspace(-50, 70, C)], 1)
    emb, W, b, labels = synth.head_inputs(B * world, D, C, seed=9)
    sl = slice(rank * B, (rank + 1) * B)
    _, gW, gb = sgo.forward_backward(emb[sl], W, b, cent, labels[sl])
    dist.all_reduce(gW)
    dist.all_reduce(gb)
    gW /= world
    gb /= world
    if rank == 0:
        _, fW, fb = sgo.forward_backward(emb, W, b, cent, labels)
        q.put((float((gW - fW).abs().max()), float((gb - fb).abs().max())))
    dist.destroy_process_group()


def _dp_chunk_worker(rank, world, port, q):
    """The module's range-by-range gradient averaging (dp_chunk_bounds + _all_reduce_avg) == one all-reduce."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from geoguessr_ai_b200.super_guessr import _all_reduce_avg, dp_chunk_bounds

    C, D = 1000, 24
    g = torch.from_numpy(np.random.default_rng(rank).standard_normal((C, D), dtype=np.float32))
    whole = g.clone()
    dist.all_reduce(whole)
    whole /= world
    for c0, c1 in dp_chunk_bounds(C, 3):
        _all_reduce_avg(g[c0:c1], None, None)
    g16 = whole.clone()
    _all_reduce_avg(g16, None, torch.bfloat16)  # opt-in compressed transfer: bf16 rounding only
    if rank == 0:
        q.put((torch.equal(g, whole), float((g16 - whole).abs().max() / whole.abs().max())))
    dist.destroy_process_group()


def test_dp_chunk_bounds_cover_every_geocell_once():
    from geoguessr_ai_b200 import dp_chunk_bounds

    for C, n in [(12647, 3), (12647, 4), (200, 3), (1000, 8), (256, 2), (257, 2)]:
        b = dp_chunk_bounds(C, n)
        assert b[0][0] == 0 and b[-1][1] == C and all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
        assert all(c0 % 256 == 0 and c1 > c0 for c0, c1 in b) and len(b) <= n
    # cfg2: every range fits one wave of the 74 CTA pairs (pair tile = 256 geocells x 256 embedding columns)
    assert all(-(-(c1 - c0) // 256) * 4 <= 74 for c0, c1 in dp_chunk_bounds(12647, 3))


def test_chunked_gradient_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_chunk_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same, err16 = q.get(timeout=120)
    for p in procs:
        p.join(60)
    assert same and err16 < 1e-2


def test_data_parallel_gradient_average_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    eW, eb = q.get(timeout=120)
    for p in procs:
        p.join(60)
    assert eW < 1e-7 and eb < 1e-7


def _bwd_pair_schedule(pair, P, T, nk):
    """Python mirror of PairSchedule (geoguessr_ai_b200/csrc/head_bwd.cu): whole rounds + stream-K tail."""
    rounds = T // P
    base = rounds * P
    total = (T - base) * nk
    u0, u1 = total * pair // P, total * (pair + 1) // P
    segs = [(i * P + pair, 0, nk) for i in range(rounds)]
    if u1 > u0:
        for t in range((u1 - 1) // nk, u0 // nk - 1, -1):
            b = t * nk
            segs.append((base + t, max(u0, b) - b, min(u1, b + nk) - b))
    lower = -1 if (u0 <= 0 or total <= 0) else (u0 * P + total - 1) // total - 1
    return segs, lower


@pytest.mark.parametrize("T,nk,P", [(200, 64, 74), (150, 2, 74), (12, 5, 12), (1, 1, 1), (7, 3, 4), (75, 1, 74),
                                    (76, 9, 74), (3, 100, 2), (396, 64, 74)])
def test_head_bwd_schedule_covers_every_unit_once(T, nk, P):
    """dW schedule: every (pair-tile, k-block) unit is computed exactly once, and a pair that starts in the middle
    of a tile finds that tile's first k-blocks parked by a LOWER pair (deadlock-free wait direction)."""
    P = min(P, T)
    cover, parks = {}, {}
    scheds = [_bwd_pair_schedule(p, P, T, nk) for p in range(P)]
    for p, (segs, _) in enumerate(scheds):
        for t, k0, k1 in segs:
            for kb in range(k0, k1):
                cover[(t, kb)] = cover.get((t, kb), 0) + 1
            if k1 < nk:
                assert p not in parks  # one parking slot per pair
                parks[p] = (t, k1)
    assert len(cover) == T * nk and set(cover.values()) == {1}
    for p, (segs, lower) in enumerate(scheds):
        for t, k0, k1 in segs:
            if k0 > 0:
                assert 0 <= lower < p and parks[lower] == (t, k0)
    work = [sum(k1 - k0 for _, k0, k1 in segs) for segs, _ in scheds]
    assert max(work) - min(work) <= 1 + (0 if T % P else 0)  # balanced to one k-block


@pytest.mark.parametrize("n,world", [(12647 * 1024 + 12648, 2), (12647 * 576 + 12648, 8), (16, 8), (4, 4), (40, 1), (1000, 4)])
def test_gradient_exchange_slices_cover_the_buffer_once(n, world):
    """gg_p2p_slice (host side of the two-shot all-reduce): the ranks' slices tile the 16-byte units of the buffer."""
    from geoguessr_ai_b200 import ops

    n -= n % 4
    edges = [ops.p2p_slice(n, world, r) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == n
    for (lo, hi), (lo2, _) in zip(edges, edges[1:]):
        assert lo <= hi == lo2 and lo % 4 == 0
    sizes = [hi - lo for lo, hi in edges]
    assert max(sizes) - min(s for s in sizes[:-1] or sizes) <= 4 * world or world == 1


def test_gradient_transport_is_decided_before_any_kernel_runs():
    """comm="auto" with a rank count the own exchange kernels do not cover (3, 5, 6, 7, > 8) must fall back to NCCL
    when the buffer is set up -- not crash in the first backward; an explicit kernel choice there is an error."""
    T = gg.SuperGuessr._dp_transport
    for world in (2, 4, 8):
        assert T("auto", world, True) == ("fused" if world == 8 else "p2p")
        assert T("p2p", world, True) == "p2p" and T("nvls", world, True) == "nvls" and T("fused", world, True) == "fused"
    for world in (3, 5, 6, 7, 16):
        assert T("auto", world, True) == "nccl"
        for comm in ("fused", "p2p", "nvls"):
            with pytest.raises(ValueError):
                T(comm, world, True)
    assert T("auto", 2, False) == "nccl"  # CPU / gloo groups
    assert T("auto", 2, True, torch.bfloat16) == "nccl"  # rounded communication is an NCCL option
    assert T("nccl", 8, True) == "nccl"


def test_sharded_adamw_ownership_partitions_the_geocells():
    """Every geocell row is owned by exactly one rank, in blocks of 128 dealt round-robin (the rule the push epilogue
    and the exchange kernel share); and the optimizer refuses a CPU model loudly (no CPU path)."""
    import torch

    from geoguessr_ai_b200.sharded_adamw import ShardedAdamW

    for C in (1, 127, 128, 129, 12647):
        for world in (1, 2, 4, 8):
            masks = torch.stack([ShardedAdamW.owned_mask(C, world, r) for r in range(world)])
            assert torch.all(masks.sum(0) == 1)
            for r in range(world):
                rows = torch.nonzero(masks[r]).flatten()
                assert torch.all((rows // 128) % world == r)
    import contextlib
    import io

    import geoguessr_ai_b200 as gg

    with contextlib.redirect_stdout(io.StringIO()):
        m = gg.SuperGuessr(None, panorama=True, embed_dim=64)
    import pytest

    with pytest.raises(RuntimeError, match="CUDA"):
        m.sharded_adamw(lr=1e-3)
