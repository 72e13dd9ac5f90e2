"""Property tests of the two host-side formats against pandas restatements of the reference's own steps:
embedding rows -> per-location records (training/load_sqlite_dataset.py:121-147: sort_values, groupby, drop rows
without a blob, first valid row's lat / lon) and proto_df rows -> per-geocell cluster order
(models/utils.py:166-168: groupby("geocell_index") + reset_index).  The fixed fixtures in tests/golden pin the same
functions to the reference executed once; these runs vary the shapes."""
import numpy as np
import pandas as pd
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from geoguessr_ai_b200 import embedding_store as es
from geoguessr_ai_b200 import proto_builder as pb

ids = st.sampled_from(["a", "B", "b", "loc_1", "loc_10", "loc_2", "zz", "Ä", "é", "0"])
row = st.tuples(ids, st.floats(-80, 80, allow_nan=False, width=32), st.floats(-180, 180, allow_nan=False, width=32),
                st.sampled_from([0, 90, 180, 270]), st.one_of(st.none(), st.binary(min_size=4, max_size=4)))


@settings(max_examples=150, deadline=None)
@given(st.lists(row, max_size=40))
def test_group_rows_equals_pandas_pipeline(rows):
    # one row per (location, heading): the table's primary key
    uniq = {}
    for r in rows:
        uniq[(r[0], r[3])] = r
    rows = list(uniq.values())
    got = list(es.group_rows(sorted(rows, key=lambda r: (r[0], r[3]))))
    want = []
    if rows:
        df = pd.DataFrame(rows, columns=["location_id", "lat", "lon", "heading", "image"])
        for loc, grp in df.sort_values(["location_id", "heading"]).groupby("location_id"):
            valid = grp[grp["image"].notna()]
            if valid.empty:
                continue
            first = valid.iloc[0]
            want.append((loc, float(first["lat"]), float(first["lon"]), valid["heading"].tolist(), valid["image"].tolist()))
    assert got == want


cluster = st.tuples(st.integers(-2, 12), st.lists(st.integers(-3, 50), max_size=5), st.floats(-180, 180, width=32),
                    st.floats(-80, 80, width=32))


@settings(max_examples=150, deadline=None)
@given(st.lists(cluster, max_size=30), st.integers(1, 10))
def test_clusters_by_cell_equals_pandas_groupby(clusters, num_cells):
    cells = [c[0] for c in clusters]
    lists = [c[1] for c in clusters]
    lng = [c[2] for c in clusters]
    lat = [c[3] for c in clusters]
    cell_off, member_off, members, coords, rows = pb.clusters_by_cell(cells, lists, lng, lat, num_cells)
    assert cell_off.shape == (num_cells + 1,) and cell_off[0] == 0 and cell_off[-1] == len(rows)
    df = pd.DataFrame({"geocell_index": cells, "row": range(len(clusters))})
    by_cell = {int(c): g.reset_index(drop=True)["row"].tolist() for c, g in df.groupby("geocell_index")} if clusters else {}
    for c in range(num_cells):
        want_rows = by_cell.get(c, [])
        lo, hi = int(cell_off[c]), int(cell_off[c + 1])
        assert rows[lo:hi].tolist() == want_rows
        for p, r in zip(range(lo, hi), want_rows):
            assert members[member_off[p]:member_off[p + 1]].tolist() == lists[r]
            assert coords[p].tolist() == [np.float32(lng[r]), np.float32(lat[r])]


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 5000), st.sampled_from([1, 2, 4, 8]))
def test_exchange_slices_partition(n4, world):
    from geoguessr_ai_b200 import ops

    edges = [ops.p2p_slice(4 * n4, world, r) for r in range(world)]
    covered = np.zeros(4 * n4, dtype=np.int32)
    for lo, hi in edges:
        assert 0 <= lo <= hi <= 4 * n4 and lo % 4 == 0 and hi % 4 == 0
        covered[lo:hi] += 1
    assert (covered == 1).all()


def test_enable_data_parallel_argument_checks():
    import torch

    import geoguessr_ai_b200 as gg

    m = gg.SuperGuessr(None, panorama=True, embed_dim=16, centroids=torch.zeros(8, 2))
    with pytest.raises(RuntimeError):
        m.enable_data_parallel()  # no process group
