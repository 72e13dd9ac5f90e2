"""Embedding-SQLite reader: grouping pinned to the reference's own loader
(tests/golden/embedding_sqlite.json = training/load_sqlite_dataset.py:104-150 executed by
oracle/make_golden_sqlite.py), schema round trip, batch iteration."""
import json
import os
import sqlite3

import numpy as np
import pytest
import torch

from geoguessr_ai_b200 import embedding_store as es

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "embedding_sqlite.json")


def _golden():
    with open(GOLDEN) as f:
        return json.load(f)


def test_grouping_matches_reference_loader():
    g = _golden()
    rows = sorted(g["rows"], key=lambda r: (r["location_id"], r["heading"]))  # the table's primary-key order
    rows = [(r["location_id"], r["lat"], r["lon"], r["heading"],
             None if r["embedding"] is None else np.asarray(r["embedding"], np.float32).tobytes()) for r in rows]
    got = list(es.group_rows(rows))
    want = g["reference_groups"]
    assert [x[0] for x in got] == [w["location_id"] for w in want]
    for (loc, lat, lon, hs, blobs), w in zip(got, want):
        assert hs == w["headings"], loc
        assert lat == w["lat"] and lon == w["lon"], loc
        assert [np.frombuffer(b, np.float32).tolist() for b in blobs] == w["embeddings"], loc


def test_sqlite_order_by_equals_reference_sort(tmp_path):
    """SQLite's BINARY collation on UTF-8 text orders location ids like pandas / Python sort code points."""
    g = _golden()
    rows = [r for r in g["rows"] if r["embedding"] is not None]
    ids = sorted({r["location_id"] for r in rows})
    path = str(tmp_path / "order.sqlite")
    conn = sqlite3.connect(path)
    conn.executescript(es.SCHEMA)
    conn.executemany("INSERT INTO samples (location_id, lat, lon, heading, embedding, embedding_dim) VALUES (?,?,?,?,?,?)",
                     [(r["location_id"], r["lat"], r["lon"], r["heading"],
                       sqlite3.Binary(np.asarray(r["embedding"], np.float32).tobytes()), g["embedding_dim"]) for r in rows])
    conn.commit()
    conn.close()
    t = es.read_embedding_sqlite(path, incomplete="mean")
    assert t.location_ids == ids
    want = {w["location_id"]: w for w in g["reference_groups"]}
    for i, loc in enumerate(t.location_ids):
        w = want[loc]
        assert t.labels[i].tolist() == [np.float32(w["lon"]), np.float32(w["lat"])]  # (lng, lat)
        for h, e in zip(w["headings"], w["embeddings"]):
            assert t.embedding[i, es.HEADINGS.index(h)].tolist() == e
        assert t.present[i].tolist() == [h in w["headings"] for h in es.HEADINGS]
        # filled slots keep the fused query equal to the mean over the views that exist
        assert torch.allclose(t.embedding[i].mean(0), torch.tensor(w["embeddings"]).mean(0), atol=1e-6)
    t = es.read_embedding_sqlite(path, incomplete="drop")
    assert t.location_ids == [w["location_id"] for w in g["reference_groups"] if len(w["headings"]) == 4]
    assert bool(t.present.all())
    with pytest.raises(ValueError):
        es.read_embedding_sqlite(path, incomplete="error")


def test_round_trip_and_batches(tmp_path):
    torch.manual_seed(0)
    N, V, D = 37, 4, 16
    emb = torch.randn(N, V, D)
    labels = torch.stack([torch.empty(N).uniform_(-180, 180), torch.empty(N).uniform_(-60, 80)], 1)
    ids = [f"id{(7 * i) % N:03d}" for i in range(N)]  # not in sorted order
    path = str(tmp_path / "rt.sqlite")
    es.write_embedding_sqlite(path, ids, emb, labels, skip=[(3, 1)])
    t = es.read_embedding_sqlite(path)  # drops the incomplete location
    keep = sorted((i for i in range(N) if i != 3), key=lambda i: ids[i])
    assert t.location_ids == [ids[i] for i in keep]
    assert torch.equal(t.embedding, emb[keep]) and torch.equal(t.labels, labels[keep])
    assert len(es.read_embedding_sqlite(path, limit=5)) == 5

    seen = []
    for e, l, idx in es.iter_batches(t, 8):
        assert torch.equal(e, t.embedding[idx]) and torch.equal(l, t.labels[idx])
        seen.append(idx)
    assert torch.equal(torch.cat(seen), torch.arange(len(t)))
    a = torch.cat([i for _, _, i in es.iter_batches(t, 8, shuffle=True, seed=3)])
    b = torch.cat([i for _, _, i in es.iter_batches(t, 8, shuffle=True, seed=3)])
    assert torch.equal(a, b) and not torch.equal(a, torch.arange(len(t))) and sorted(a.tolist()) == list(range(len(t)))
    assert sum(i.numel() for _, _, i in es.iter_batches(t, 8, drop_last=True)) == len(t) // 8 * 8


def test_bad_blob_size_is_reported(tmp_path):
    path = str(tmp_path / "bad.sqlite")
    conn = sqlite3.connect(path)
    conn.executescript(es.SCHEMA)
    conn.execute("INSERT INTO samples (location_id, lat, lon, heading, embedding, embedding_dim) VALUES ('a',0,0,0,?,8)",
                 (sqlite3.Binary(b"\x00" * 12),))
    conn.commit()
    conn.close()
    with pytest.raises(ValueError, match="12 bytes"):
        es.read_embedding_sqlite(path)


@pytest.mark.gpu
def test_device_batches_feed_the_head(tmp_path):
    """Device batches equal the host table, and a training step from them runs through the CUDA path."""
    import geoguessr_ai_b200 as gg
    from geoguessr_ai_b200.geocells import load_packaged_centroids

    torch.manual_seed(1)
    N, V, D = 200, 4, 64
    emb = torch.randn(N, V, D)
    labels = torch.stack([torch.empty(N).uniform_(-180, 180), torch.empty(N).uniform_(-60, 80)], 1)
    path = str(tmp_path / "dev.sqlite")
    es.write_embedding_sqlite(path, [f"l{i:04d}" for i in range(N)], emb, labels)
    t = es.read_embedding_sqlite(path, pin_memory=True)
    dev = torch.device("cuda:0")
    model = gg.SuperGuessr(None, panorama=True, should_smooth_labels=True, embed_dim=D,
                           centroids=load_packaged_centroids()).to(dev).train()
    n = 0
    for e, l, idx in es.iter_batches(t, 64, device=dev, shuffle=True, seed=5):
        assert e.is_cuda and torch.equal(e.cpu(), t.embedding[idx]) and torch.equal(l.cpu(), t.labels[idx])
        out = model(embedding=e, labels=l, labels_clf=torch.zeros(e.shape[0], dtype=torch.int64, device=dev))
        out.loss.backward()
        assert torch.isfinite(out.loss)
        n += idx.numel()
    assert n == N
