"""ctypes binding of libgeoguessr_b200.so (the C ABI declared in include/geoguessr_b200.h).

No torch types cross this boundary: tensors are passed as raw device pointers
(``tensor.data_ptr()``) plus sizes, the stream as ``torch.cuda.current_stream().cuda_stream``.
There is no fallback: if the library is missing or a launcher fails, the call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgeoguessr_b200.so")
ABI_VERSION = 4

P, I, L, F = c_void_p, c_int, c_longlong, c_float

# name -> (restype, argtypes); mirrors include/geoguessr_b200.h one to one
SIGNATURES = {
    "gg_abi_version": (I, []),
    "gg_last_error": (c_char_p, []),
    "gg_head_logits_ld": (I, [I]),
    "gg_head_bias_pad": (I, [I]),
    "gg_hav_cpad": (I, [I]),
    "gg_centroid_table_floats": (c_size_t, [I]),
    "gg_centroid_table_workspace_bytes": (c_size_t, [I]),
    "gg_hav_row_stats_bytes": (c_size_t, [I, I]),
    "gg_head_fwd_workspace_bytes": (c_size_t, [I, I, I]),
    "gg_head_fwd_ticket_bytes": (c_size_t, [I]),
    "gg_debug_head_fwd_timeline": (None, [P]),
    "gg_debug_head_bwd_timeline": (None, [P]),
    "gg_debug_nvlink_store_probe": (I, [P, c_size_t, I, I, I, P]),
    "gg_head_bwd_workspace_bytes": (c_size_t, [I]),
    "gg_hav_ce_workspace_bytes": (c_size_t, [I, I]),
    "gg_hav_ce_db_parts": (I, [I, I]),
    "gg_proto_retrieve_workspace_bytes": (c_size_t, [I, I, I, I]),
    "gg_fuse_headings": (I, [P, I, P, I, I, I, I, P, P]),
    "gg_prepare_head_weights": (I, [P, P, P, P, I, I, I, P]),
    "gg_fuse_and_prepare": (I, [P, I, P, I, I, I, P, P, P, P, I, I, P]),
    "gg_cast_bf16": (I, [P, P, L, I, I, P]),
    "gg_row_sqnorm_bf16": (I, [P, L, I, I, P, P]),
    "gg_head_fwd": (I, [P, P, P, I, I, I, P, I, I, P, P, P, P, P, P, P, P, P]),
    "gg_centroid_unit_vectors": (I, [P, P, I, P, P]),
    "gg_hav_row_stats": (I, [P, P, I, I, F, F, P, P, P, P]),
    "gg_hav_ce_fwd_bwd": (I, [P, I, P, P, P, I, I, F, P, P, P, P, P, F, P]),
    "gg_hard_ce_fwd_bwd": (I, [P, I, P, P, I, I, P, P, P]),
    "gg_loss_mean": (I, [P, I, F, P, P]),
    "gg_head_bwd": (I, [P, I, P, I, I, I, I, F, P, P, P, P, I, I, P, P, I, I, I, P]),
    "gg_head_dx": (I, [P, I, P, I, I, I, I, F, P, I, P, P]),
    "gg_topk_accuracy": (I, [P, I, P, I, P, P]),
    "gg_split3_bf16": (I, [P, L, I, I, P, I, P, P]),
    "gg_linear_bf16": (I, [P, I, P, I, P, I, I, I, P, I, P]),
    "gg_hier_attention": (I, [P, I, I, I, I, P, P]),
    "gg_proto_take_image_coords": (I, [P, P, L, P]),
    "gg_proto_record_ids": (I, [P, L, P, P]),
    "gg_proto_group_cells": (I, [P, I, P]),
    "gg_proto_retrieve": (I, [P, P, I, I, P, I, I, P, P, P, L, P, I, I, P, I, I, I, I, P, P, P]),
    "gg_proto_refine": (I, [P, I, L, P, I, P, I, P, I, I, F, F, P, P, P, P, P, P]),
    "gg_build_prototypes": (I, [P, L, I, I, P, P, P, L, P, P, P, P]),
    "gg_p2p_slice": (None, [c_size_t, I, I, P, P]),
    "gg_p2p_allreduce_avg": (I, [P, I, I, c_size_t, P]),
    "gg_nvls_allreduce_avg": (I, [P, I, I, c_size_t, P]),
    "gg_grad_ctrl_bytes": (c_size_t, []),
    "gg_grad_stage_floats": (c_size_t, [I, I, I]),
    "gg_grad_exchange": (I, [P, P, P, P, P, I, I, I, I, I, P]),
    "gg_grad_exchange_adamw": (I, [P, P, P, P, P, P, P, I, I, I, I, P, P, P, P, P, P, P, P, I, P]),
}

_lib = None


class GeoguessrB200Error(RuntimeError):
    pass


def load():
    """Load the CUDA library (once).  Raises if it has not been built -- there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GeoguessrB200Error(
            f"{LIB_PATH} not found: the sm_100a CUDA library is not built. Run "
            "`python -m geoguessr_ai_b200.build` (needs nvcc); there is no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.gg_abi_version() != ABI_VERSION:
        raise GeoguessrB200Error(
            f"ABI mismatch: library reports {lib.gg_abi_version()}, binding expects {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().gg_last_error().decode("utf-8", "replace")
        raise GeoguessrB200Error(f"{what} failed (code {rc}): {msg}")
