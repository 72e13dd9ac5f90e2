"""Prototype bank builder: the reference's `ProtoRefiner(protos=None)` offline job with the vision
encoder replaced by stored embeddings.

Reference flow (models/proto_refiner.py:89-103, :313-345): for every geocell, `ProtoDataManager`
(models/utils.py:98-181) returns the rows of `proto_df.csv` (data/geocells/geocell_manager.py:112-136: one row per
cluster with `geocell_index`, `indices` = member location indices, `centroid_lat` / `centroid_lng` = the CELL's
geometric centroid); every row becomes one prototype = running fp32 mean over the members of the member's mean over
its headings (`Embeddings.generate_embeddings`, :461-517).  At query time the prototype's coordinates are the row's
(centroid_lng, centroid_lat) (`_within_cluster_refinement`, :251-252 -- the only executable branch).

Here the member embeddings come from an (L, V, D) fp32 array (e.g. `embedding_store.read_embedding_sqlite`) and the
means run in one HBM-bound kernel (gg_build_prototypes); the result is the CSR triple `ProtoRefiner(bank=...)`
takes.  The host side only parses and orders the table; no arithmetic of the path runs in Python.
"""
from __future__ import annotations

import ast
from typing import Iterable, List, Sequence

import numpy as np
import torch

from . import _lib, ops


def parse_indices(val) -> List[int]:
    """The `indices` column in any of the forms `ProtoDataManager._parse_indices_value` accepts
    (models/utils.py:121-157): list / tuple, a literal string "[1, 2]" / "(1,2)" / "7", a loose "1, 2" string,
    NaN / "" -> empty; entries that are not integers are dropped."""
    if isinstance(val, (list, tuple, np.ndarray)):
        cand = list(val)
    elif val is None or (isinstance(val, float) and np.isnan(val)):
        cand = []
    elif isinstance(val, str):
        s = val.strip()
        if s == "":
            cand = []
        else:
            try:
                obj = ast.literal_eval(s)
            except Exception:  # noqa: BLE001
                obj = [part for part in s.strip("[](){}").split(",") if part != ""]
            cand = list(obj) if isinstance(obj, (list, tuple)) else [obj]
    else:
        cand = [val]
    out = []
    for x in cand:
        try:
            out.append(int(x))
        except Exception:  # noqa: BLE001
            try:
                out.append(int(str(x).strip()))
            except Exception:  # noqa: BLE001
                continue
    return out


def clusters_by_cell(geocell_index: Sequence[int], indices: Iterable, centroid_lng: Sequence[float],
                     centroid_lat: Sequence[float], num_cells: int):
    """Order the clusters as the reference serves them: grouped by geocell (models/utils.py:166-168, pandas
    groupby: ascending cell, table order inside a cell).  Returns (cell_off (C+1) int32, member_off (P+1) int64,
    members (M) int32, coords (P,2) fp32 (lng, lat), row (P) int64 = source row of every prototype).  Rows whose
    geocell_index is outside [0, num_cells) are left out (no geocell could ever ask for them)."""
    cells = np.asarray([int(c) for c in geocell_index], dtype=np.int64)
    rows = np.flatnonzero((cells >= 0) & (cells < num_cells))
    rows = rows[np.argsort(cells[rows], kind="stable")]
    lists = [parse_indices(v) for v in indices]
    cell_off = np.zeros(num_cells + 1, dtype=np.int64)
    np.add.at(cell_off, cells[rows] + 1, 1)
    cell_off = np.cumsum(cell_off)
    member_off = np.zeros(len(rows) + 1, dtype=np.int64)
    member_off[1:] = np.cumsum([len(lists[r]) for r in rows])
    # location indices beyond int32 cannot exist (they index the embedding array); anything else is "out of range"
    members = np.asarray([min(max(i, -2**31), 2**31 - 1) for r in rows for i in lists[r]], dtype=np.int32)
    lng = np.asarray(centroid_lng, dtype=np.float32)
    lat = np.asarray(centroid_lat, dtype=np.float32)
    coords = np.stack([lng[rows], lat[rows]], 1) if len(rows) else np.zeros((0, 2), np.float32)
    return cell_off.astype(np.int32), member_off, members, coords.astype(np.float32), rows


def build_prototype_bank(embedding: torch.Tensor, proto_df, num_cells: int, valid=None, device="cuda",
                         return_f32: bool = False):
    """(cell_off, bank (P,D) bf16, coords (P,2) fp32) -- the ``bank=`` argument of ProtoRefiner -- from location
    embeddings (L, V, D) or (L, D) fp32 and the reference's proto table.

    proto_df: a pandas DataFrame (or any mapping of columns) with `geocell_index`, `indices`, `centroid_lng`,
        `centroid_lat` (data/geocells/proto_df.csv).
    valid: (L,) bool, False for locations the reference skips because their coordinates are not finite
        (proto_refiner.py:470-472); None = all valid.
    return_f32: also return the unrounded fp32 means and the per-prototype member counts."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.GeoguessrB200Error("build_prototype_bank runs on the sm_100a path only; there is no CPU fallback")
    emb = embedding if embedding.dim() == 3 else embedding.unsqueeze(1)
    emb = emb.to(dev, torch.float32).contiguous()
    L, V, D = emb.shape
    cell_off, member_off, members, coords, _ = clusters_by_cell(
        proto_df["geocell_index"], proto_df["indices"], proto_df["centroid_lng"], proto_df["centroid_lat"], num_cells)
    P = len(member_off) - 1
    bank = torch.empty((P, D), dtype=torch.bfloat16, device=dev)
    f32 = torch.empty((P, D), dtype=torch.float32, device=dev) if return_f32 else None
    cnt = torch.empty((P,), dtype=torch.int32, device=dev) if return_f32 else None
    moff = torch.from_numpy(member_off).to(dev)
    mem = torch.from_numpy(members).to(dev) if len(members) else torch.zeros(1, dtype=torch.int32, device=dev)
    val = None if valid is None else torch.as_tensor(valid).to(dev, torch.uint8).contiguous()
    if val is not None and val.numel() != L:
        raise ValueError("valid must have one entry per location")
    dev = ops._need_cuda(emb, moff, mem)
    ops._call("gg_build_prototypes", _lib.load().gg_build_prototypes, dev, emb.data_ptr(), L, V, D, moff.data_ptr(), mem.data_ptr(),
              0 if val is None else val.data_ptr(), P, bank.data_ptr(), 0 if f32 is None else f32.data_ptr(),
              0 if cnt is None else cnt.data_ptr(), ops._stream(dev))
    out = (torch.from_numpy(cell_off), bank, torch.from_numpy(coords))
    return out + (f32, cnt) if return_f32 else out
