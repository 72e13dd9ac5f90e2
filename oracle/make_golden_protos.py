#!/usr/bin/env python
"""Golden vectors for the prototype table parsing / grouping, produced by EXECUTING the reference's
`models.utils.ProtoDataManager` (models/utils.py:98-181) on a synthetic proto_df in the shape
data/geocells/geocell_manager.py:112-136 writes, round-tripped through CSV as the reference does.

    python oracle/make_golden_protos.py        (needs /root/reference; build container only)
"""
import io
import json
import os
import sys

import numpy as np
import pandas as pd

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "proto_table.json")


def main():
    import types

    cfg = types.ModuleType("config")  # config.py builds HF TrainingArguments at import (needs accelerate): stubbed
    cfg.LABEL_SMOOTHING_CONSTANT = 65
    cfg.__getattr__ = lambda name: "unused-" + name
    sys.modules["config"] = cfg
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    import transformers  # noqa: F401  (models.utils imports Trainer)
    from models.utils import ProtoDataManager  # the reference, unmodified

    rng = np.random.default_rng(11)
    C = 12
    rows = []
    order = [3, 0, 7, 3, 11, 0, 3, 9, 7, 5, 0]  # clusters arrive cell-interleaved; cells 1, 2, 4, 6, 8, 10 have none
    for k, cell in enumerate(order):
        n = int(rng.integers(0, 6))
        idx = [int(i) for i in rng.integers(-2, 45, size=n)]  # a few out of range on purpose
        rows.append(dict(geocell_index=cell, country="XX", admin1="YY", cell_id=f"c{cell}", cluster_id=k, count=n,
                         indices=idx, centroid_lat=float(rng.uniform(-60, 80)), centroid_lng=float(rng.uniform(-180, 180))))
    df = pd.DataFrame(rows)
    csv = df.to_csv(index=False)               # geocell_manager.py:136 writes CSV: lists become strings
    df2 = pd.read_csv(io.StringIO(csv))
    # hand-edited / legacy spellings the reference parser also accepts (models/utils.py:121-157)
    df2.loc[1, "indices"] = "(4, 5, 6)"
    df2.loc[2, "indices"] = "7"
    df2.loc[4, "indices"] = " 1, 2 ,x, 3 "
    df2.loc[5, "indices"] = float("nan")
    mgr = ProtoDataManager(df2)
    cells = {}
    for c in range(C):
        g = mgr.get_indices_for_cell(c)
        cells[str(c)] = [dict(indices=[int(i) for i in r["indices"]], centroid_lng=float(r["centroid_lng"]),
                              centroid_lat=float(r["centroid_lat"]), cluster_id=int(r["cluster_id"])) for _, r in g.iterrows()]
    table = [dict(geocell_index=int(r["geocell_index"]), indices=(None if isinstance(r["indices"], float) else r["indices"]),
                  centroid_lng=float(r["centroid_lng"]), centroid_lat=float(r["centroid_lat"]), cluster_id=int(r["cluster_id"]))
             for _, r in df2.iterrows()]
    with open(OUT, "w") as f:
        json.dump(dict(num_cells=C, table=table, reference_cells=cells,
                       source="models/utils.py:98-181 ProtoDataManager executed on this table"), f, indent=1)
    print(f"wrote {OUT}: {len(table)} clusters over {C} cells")


if __name__ == "__main__":
    main()
