#!/usr/bin/env python
"""Golden vectors for the embedding-SQLite reader, produced by EXECUTING the reference.

The reference groups the per-image rows of the `samples` table into one record per location with
training/load_sqlite_dataset.py:104-150 (`_build_panorama_dataframe`: sort by (location_id, heading), drop rows
whose blob is missing, lat/lon from the first valid row).  This script builds a small synthetic table in the
embedding schema of backend/s3bucket.py:848-859 (blob = fp32[embedding_dim].tobytes(), :946), runs the
reference function on it (its only non-stdlib import, backend.s3bucket, is stubbed: it is S3 plumbing) and
writes tests/golden/embedding_sqlite.json: the rows as inserted and the grouping the reference produced.

    python oracle/make_golden_sqlite.py        (needs /root/reference; run in the build container only)
"""
import json
import os
import sys
import types

import numpy as np
import pandas as pd

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "embedding_sqlite.json")


def main():
    stub = types.ModuleType("backend.s3bucket")
    stub.BUCKET = stub.DATASET_SQLITE_PREFIX = ""
    stub.get_json = stub.s3 = None
    pkg = types.ModuleType("backend")
    pkg.s3bucket = stub
    sys.modules["backend"], sys.modules["backend.s3bucket"] = pkg, stub
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    from training.load_sqlite_dataset import _build_panorama_dataframe  # the reference, unmodified

    rng = np.random.default_rng(7)
    D = 8
    rows = []
    locs = ["loc_0007", "loc_0001", "Zeta", "alpha", "loc_0010", "loc_0002", "ümlaut", "loc_0003", "solo", "empty"]
    for li, loc in enumerate(locs):
        lat, lon = float(rng.uniform(-60, 80)), float(rng.uniform(-180, 180))
        headings = [0, 90, 180, 270]
        if loc == "loc_0002":
            headings = [270, 0]          # incomplete location
        if loc == "solo":
            headings = [180]
        rng.shuffle(headings)
        for h in headings:
            emb = rng.standard_normal(D).astype(np.float32)
            missing = (loc == "empty") or (loc == "loc_0003" and h == 90)  # NULL blobs are dropped by the reference
            rows.append(dict(location_id=loc, lat=lat + (0.001 * h if loc == "Zeta" else 0.0), lon=lon, heading=int(h),
                             embedding=None if missing else emb.tolist()))
    order = rng.permutation(len(rows))
    rows = [rows[i] for i in order]

    df = pd.DataFrame({
        "location_id": [r["location_id"] for r in rows], "lat": [r["lat"] for r in rows], "lon": [r["lon"] for r in rows],
        "heading": [r["heading"] for r in rows],
        # the reference function reads the blob column under the name "image"
        "image": [None if r["embedding"] is None else np.asarray(r["embedding"], np.float32).tobytes() for r in rows],
    })
    pano = _build_panorama_dataframe(df)
    groups = []
    for _, g in pano.iterrows():
        groups.append(dict(location_id=g["location_id"], lat=float(g["lat"]), lon=float(g["lon"]),
                           headings=[int(h) for h in g["headings"]],
                           embeddings=[np.frombuffer(b, np.float32).tolist() for b in g["images"]]))
    with open(OUT, "w") as f:
        json.dump(dict(embedding_dim=D, rows=rows, reference_groups=groups,
                       source="training/load_sqlite_dataset.py:104-150 executed on these rows"), f, indent=1)
    print(f"wrote {OUT}: {len(rows)} rows -> {len(groups)} locations")


if __name__ == "__main__":
    main()
