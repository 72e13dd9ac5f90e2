"""CPU: the C-ABI library loads without a GPU and exports every symbol include/geoguessr_b200.h declares."""
import ctypes
import os
import re

import pytest
import torch

from conftest import REPO
from geoguessr_ai_b200 import _lib, ops


def header_symbols():
    src = open(os.path.join(REPO, "include", "geoguessr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gg_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_header_symbols():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m geoguessr_ai_b200.build` (or __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes binding and header disagree"
    assert _lib.load().gg_abi_version() == _lib.ABI_VERSION


def test_layout_helpers():
    lib = _lib.load()
    assert lib.gg_head_logits_ld(12647) == 12800 and lib.gg_head_logits_ld(64) == 256
    assert lib.gg_head_bias_pad(12647) == 12800
    assert lib.gg_hav_cpad(12647) == 12800
    assert lib.gg_centroid_table_floats(12647) == 3 * 12800 + 4 * 12672 + 4 * 198
    assert lib.gg_head_fwd_workspace_bytes(4096, 12647, 5) >= 148 * 2 * 128 * 12 * 4
    assert lib.gg_proto_retrieve_workspace_bytes(64, 5, 1024, 12647) > 64 * 5 * 1024 * 2


def test_argument_errors_are_reported_not_crashed():
    lib = _lib.load()
    rc = lib.gg_head_fwd(0, 0, 0, 16, 100, 12, 0, 0, 5, 0, 0, 0, 0, 0, 0, 0, 0, 0)  # D % 8 != 0
    assert rc == 1 and b"multiple of 8" in lib.gg_last_error()
    rc = lib.gg_hav_ce_fwd_bwd(0, 0, 0, 0, 0, 0, 10, 65.0, 0, 0, 0, 0, 0, 1.0, 0)
    assert rc == 1


def test_no_cpu_fallback():
    with pytest.raises(_lib.GeoguessrB200Error):
        ops.fuse_headings(torch.zeros(2, 4, 8))
    import geoguessr_ai_b200 as gg

    m = gg.SuperGuessr(None, panorama=True, embed_dim=64, centroids=torch.zeros(10, 2), serving=True).eval()
    with pytest.raises(_lib.GeoguessrB200Error):
        m(embedding=torch.zeros(2, 4, 64))
