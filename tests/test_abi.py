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


def test_every_launcher_has_a_python_caller():
    """Each gg_* entry point of the binding is used by the package itself (ops.py / proto_builder.py / the modules):
    a wrapper lost in an edit shows up here, on the CPU, not as an AttributeError on the GPU box."""
    pkg = os.path.join(REPO, "geoguessr_ai_b200")
    text = ""
    for name in os.listdir(pkg):
        if name.endswith(".py") and name != "_lib.py":
            text += open(os.path.join(pkg, name)).read()
    for tool in ("head_fwd_timeline.py", "head_bwd_timeline.py", "symm_probe.py"):
        text += open(os.path.join(REPO, "tools", tool)).read()
    missing = [s for s in _lib.SIGNATURES if s not in text and s not in ("gg_abi_version", "gg_last_error")]
    assert not missing, f"no Python caller for {missing}"
    for fn in ("head_dx", "topk_accuracy", "split3_bf16", "linear_bf16", "hier_fuse", "grad_exchange", "grad_exchange_adamw",
               "grad_stage_floats",
               "proto_record_ids", "proto_take_image_coords", "cast_bank_bf16", "proto_group_cells", "head_forward",
               "head_backward", "hav_ce", "hard_ce", "fuse_headings", "fuse_and_prepare", "proto_retrieve", "proto_refine"):
        assert hasattr(ops, fn), fn
