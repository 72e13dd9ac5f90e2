"""Embedding-SQLite reader: the on-disk format in front of the head-training path.

The reference stores pre-computed encoder embeddings one row per (location, heading) in a SQLite table
(backend/s3bucket.py:848-859 CLIP, :1160-1170 TinyViT):

    samples(location_id TEXT, lat REAL, lon REAL, heading INTEGER, capture_date, pano_id, batch_date,
            embedding BLOB  -- fp32[embedding_dim], `emb.numpy().tobytes()` (:946)
            embedding_dim INTEGER, PRIMARY KEY (location_id, heading)) WITHOUT ROWID

and groups per-image rows into one record per location with training/load_sqlite_dataset.py:104-150: sort by
(location_id, heading), drop rows without a blob, skip locations left empty, lat / lon from the first valid row.
This module reads that table into the (N, 4, D) fp32 layout `SuperGuessr.forward(embedding=...)` takes
(models/super_guessr.py:336-347), labels as (lng, lat) (main_coordinator_idun_s3.py:388), and feeds device
batches through pinned, double-buffered host-to-device copies -- the same pipeline `bench.py` times as `e2e`.
Host-side only: no arithmetic of the path runs here (the heading mean stays in gg_fuse_headings).
"""
from __future__ import annotations

import sqlite3
from dataclasses import dataclass
from pathlib import Path
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

HEADINGS = (0, 90, 180, 270)  # training/load_sqlite_dataset.py:122, backend/s3bucket.py:63

SCHEMA = """
CREATE TABLE IF NOT EXISTS samples (
  location_id TEXT NOT NULL,
  lat REAL NOT NULL,
  lon REAL NOT NULL,
  heading INTEGER NOT NULL,
  capture_date TEXT,
  pano_id TEXT,
  batch_date TEXT,
  embedding BLOB NOT NULL,
  embedding_dim INTEGER NOT NULL,
  PRIMARY KEY (location_id, heading)
) WITHOUT ROWID;
"""


@dataclass
class EmbeddingTable:
    """One record per location, in the reference's order (sorted by location_id)."""
    location_ids: List[str]
    embedding: torch.Tensor  # (N, V, D) fp32, heading slots in the order of `headings`
    labels: torch.Tensor     # (N, 2) fp32 (lng, lat) -- the coordinate order of the whole path
    present: torch.Tensor    # (N, V) bool: False where the slot was filled (incomplete="mean")
    headings: Tuple[int, ...]

    def __len__(self) -> int:
        return len(self.location_ids)


def _connect_readonly(path: str) -> sqlite3.Connection:
    # strictly read-only, like training/load_sqlite_dataset.py:54-61 (no -wal / -shm files next to the dataset)
    conn = sqlite3.connect(f"{Path(path).resolve().as_uri()}?mode=ro", uri=True)
    conn.execute("PRAGMA query_only = 1")
    return conn


def group_rows(rows, headings: Sequence[int] = HEADINGS):
    """The reference's grouping (load_sqlite_dataset.py:121-147) on rows (location_id, lat, lon, heading, blob)
    already sorted by (location_id, heading): yields (location_id, lat, lon, [heading...], [blob...]) per location
    with at least one blob; lat / lon come from the first row that has one."""
    cur_id, lat, lon, hs, blobs = None, 0.0, 0.0, [], []
    for loc, la, lo, h, blob in rows:
        if loc != cur_id:
            if cur_id is not None and blobs:
                yield cur_id, lat, lon, hs, blobs
            cur_id, hs, blobs = loc, [], []
        if blob is None:
            continue
        if not blobs:
            lat, lon = float(la), float(lo)
        hs.append(int(h))
        blobs.append(bytes(blob) if isinstance(blob, memoryview) else blob)
    if cur_id is not None and blobs:
        yield cur_id, lat, lon, hs, blobs


def read_embedding_sqlite(path: str, headings: Sequence[int] = HEADINGS, incomplete: str = "drop",
                          pin_memory: bool = False, limit: Optional[int] = None) -> EmbeddingTable:
    """Read the whole `samples` table into an EmbeddingTable.

    incomplete: what to do with a location that lacks one of ``headings`` --
      "drop"  leave it out (fixed (N, V, D) batches of complete panoramas),
      "mean"  fill the missing slots with the mean of the present ones, so the fused query (mean over the V
              slots, super_guessr.py:347) equals the mean over the views that exist -- what the reference
              computes for a shorter panorama,
      "error" raise.
    Rows whose heading is not in ``headings`` are ignored.  limit: stop after that many locations."""
    if incomplete not in ("drop", "mean", "error"):
        raise ValueError("incomplete must be 'drop', 'mean' or 'error'")
    headings = tuple(int(h) for h in headings)
    slot = {h: i for i, h in enumerate(headings)}
    V = len(headings)
    conn = _connect_readonly(path)
    try:
        cur = conn.execute("SELECT location_id, lat, lon, heading, embedding, embedding_dim FROM samples "
                           "ORDER BY location_id, heading")  # the table's own (primary key) order
        dims = set()

        def rows():
            for loc, la, lo, h, blob, dim in cur:
                if blob is not None:
                    dims.add(int(dim))
                    if len(blob) != 4 * int(dim):
                        raise ValueError(f"{path}: embedding of ({loc!r}, {h}) has {len(blob)} bytes, embedding_dim says "
                                         f"{dim} fp32 values")
                yield loc, la, lo, h, blob

        ids, embs, labels, present = [], [], [], []
        for loc, lat, lon, hs, blobs in group_rows(rows(), headings):
            if len(dims) > 1:
                raise ValueError(f"{path}: mixed embedding_dim values {sorted(dims)}")
            D = next(iter(dims))
            e = np.zeros((V, D), dtype=np.float32)
            have = np.zeros(V, dtype=bool)
            for h, blob in zip(hs, blobs):
                if h in slot:
                    e[slot[h]] = np.frombuffer(blob, dtype=np.float32, count=D)
                    have[slot[h]] = True
            if not have.any():
                continue
            if not have.all():
                if incomplete == "error":
                    raise ValueError(f"{path}: location {loc!r} has headings {hs}, wanted {list(headings)}")
                if incomplete == "drop":
                    continue
                e[~have] = e[have].mean(axis=0, dtype=np.float32)
            ids.append(loc)
            embs.append(e)
            labels.append((lon, lat))
            present.append(have)
            if limit is not None and len(ids) >= limit:
                break
    finally:
        conn.close()
    if not ids:
        raise ValueError(f"{path}: no location with usable embeddings")  # load_sqlite_dataset.py:149-150
    emb = torch.from_numpy(np.stack(embs))
    lab = torch.tensor(labels, dtype=torch.float32)
    if pin_memory:
        emb, lab = emb.pin_memory(), lab.pin_memory()
    return EmbeddingTable(ids, emb, lab, torch.from_numpy(np.stack(present)), headings)


def write_embedding_sqlite(path: str, location_ids: Sequence[str], embedding, labels, headings: Sequence[int] = HEADINGS,
                           skip: Sequence[Tuple[int, int]] = ()):
    """Write (N, V, D) embeddings + (N, 2) (lng, lat) labels in the reference's schema (s3bucket.py:848-859,
    blob = fp32 bytes :946).  skip: (location index, heading slot) pairs to leave out.  For fixtures / synthetic
    data -- the reference's own writer needs the encoder and S3."""
    emb = np.ascontiguousarray(torch.as_tensor(embedding).detach().cpu().numpy(), dtype=np.float32)
    lab = torch.as_tensor(labels).detach().cpu().numpy()
    N, V, D = emb.shape
    assert len(location_ids) == N and V == len(headings)
    skip = set(skip)
    conn = sqlite3.connect(path)
    try:
        conn.executescript(SCHEMA)
        conn.executemany(
            "INSERT OR REPLACE INTO samples (location_id, lat, lon, heading, capture_date, pano_id, batch_date, embedding, "
            "embedding_dim) VALUES (?, ?, ?, ?, NULL, NULL, NULL, ?, ?)",
            [(location_ids[i], float(lab[i, 1]), float(lab[i, 0]), int(headings[v]), sqlite3.Binary(emb[i, v].tobytes()), D)
             for i in range(N) for v in range(V) if (i, v) not in skip])
        conn.commit()
    finally:
        conn.close()


def iter_batches(table: EmbeddingTable, batch_size: int, device=None, shuffle: bool = False, seed: int = 0,
                 drop_last: bool = False) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
    """Yield (embedding (b, V, D), labels (b, 2), index (b,)) batches on ``device``.

    CUDA device: every batch is gathered into one of two pinned staging buffers and copied on a private stream
    while the previous batch is being consumed; the consumer's stream is made to wait for the copy (and the
    staging buffer is not reused before its copy has completed), so the step after ``next()`` never waits for
    PCIe unless the link is the bottleneck."""
    N = len(table)
    order = torch.randperm(N, generator=torch.Generator().manual_seed(seed)) if shuffle else torch.arange(N)
    starts = list(range(0, N - (N % batch_size if drop_last else 0), batch_size))
    dev = torch.device(device) if device is not None else torch.device("cpu")
    if dev.type != "cuda":
        for s in starts:
            idx = order[s:s + batch_size]
            yield table.embedding[idx], table.labels[idx], idx
        return
    V, D = table.embedding.shape[1:]
    copy = torch.cuda.Stream(device=dev)
    stage = [(torch.empty((batch_size, V, D), dtype=torch.float32).pin_memory(),
              torch.empty((batch_size, 2), dtype=torch.float32).pin_memory()) for _ in range(2)]
    copied = [torch.cuda.Event(), torch.cuda.Event()]

    def issue(k, s):
        idx = order[s:s + batch_size]
        b = idx.numel()
        copied[k].synchronize()  # the staging buffer's previous copy has left the host
        torch.index_select(table.embedding, 0, idx, out=stage[k][0][:b])
        torch.index_select(table.labels, 0, idx, out=stage[k][1][:b])
        with torch.cuda.stream(copy):
            e = stage[k][0][:b].to(dev, non_blocking=True)
            l = stage[k][1][:b].to(dev, non_blocking=True)
            copied[k].record(copy)
        return e, l, idx

    pending = issue(0, starts[0]) if starts else None
    for n, s in enumerate(starts):
        e, l, idx = pending
        ev = copied[n % 2]
        pending = issue((n + 1) % 2, starts[n + 1]) if n + 1 < len(starts) else None
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ev)
        e.record_stream(cur)
        l.record_stream(cur)
        yield e, l, idx
