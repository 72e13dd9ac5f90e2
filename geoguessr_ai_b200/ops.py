"""Thin tensor-level wrappers over the C ABI (include/geoguessr_b200.h).

Each function allocates its outputs with torch (device memory + caching allocator are torch's
job -- plumbing), passes raw device pointers and the current CUDA stream to one launcher of
libgeoguessr_b200.so, and returns.  No arithmetic on the hot path happens in PyTorch here, and
nothing here works without a CUDA device: CPU tensors are rejected.
"""
from __future__ import annotations

import math

import torch

from . import _lib

FAR_KM_DEFAULT = 65.0 * math.log(2.0 ** 32)  # ~1442 km: targets below 2^-32 of the nearest cell's are dropped


# Optional per-launcher timing (bench.py): when enabled, every C-ABI call is bracketed by CUDA events
# recorded on the launching stream (the current torch stream), name -> [(start, end), ...].
_events = None


def enable_timing(on: bool = True):
    global _events
    _events = {} if on else None


def timing_enabled() -> bool:
    return _events is not None


def timing_ms():
    """Synchronises and returns {launcher: [ms per call, ...]} for the calls since enable_timing()."""
    torch.cuda.synchronize()
    return {k: [a.elapsed_time(b) for a, b in v] for k, v in (_events or {}).items()}


def _call(name, fn, dev, *args):
    """One launcher call on device ``dev``: every launcher works on the CURRENT CUDA device (its stream argument,
    cudaFuncSetAttribute, the SM count), so the tensors' device is made current around the call -- a module on
    cuda:1 while cuda:0 is current would otherwise launch on GPU 0 with GPU-1 pointers."""
    if dev is not None and dev.index is not None and torch.cuda.current_device() != dev.index:
        with torch.cuda.device(dev):
            return _call(name, fn, None, *args)
    if _events is None:
        _lib.check(fn(*args), name)
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(torch.cuda.current_stream())
    _lib.check(fn(*args), name)
    b.record(torch.cuda.current_stream())
    _events.setdefault(name, []).append((a, b))


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream(dev=None):
    return torch.cuda.current_stream(dev).cuda_stream


def _need_cuda(*tensors):
    """All given tensors (None entries skipped) must live on ONE CUDA device; returns it."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.GeoguessrB200Error(
                "geoguessr_ai_b200 kernels need CUDA tensors (sm_100a); got a CPU tensor and there is no CPU fallback")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise _lib.GeoguessrB200Error(
                f"geoguessr_ai_b200 kernels need all tensors of a call on one device; got {dev} and {t.device}")
    return dev


_IN_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}  # GG_IN_F32 / GG_IN_BF16 / GG_IN_F16


def _as_embedding(t):
    """Embeddings are read by the fusion kernel in their own dtype: fp32 (the reference's storage format,
    backend/s3bucket.py:848-859) or, opt-in, bf16 / fp16 (half the bytes over PCIe).  Anything else is rejected
    rather than cast by an eager torch op: no arithmetic of the path runs in PyTorch."""
    code = _IN_DTYPES.get(t.dtype)
    if code is None:
        raise _lib.GeoguessrB200Error(f"embeddings must be float32, bfloat16 or float16 (got {t.dtype})")
    return (t if t.is_contiguous() else t.contiguous()), code


_ticket_cache = {}


def _tickets(nbytes, device):
    """Zero-initialised ticket words of a launcher that counts its own CTAs' arrivals (gg_head_fwd): zeroed once,
    left zeroed by every launch, one launch at a time -- so one buffer per (device, stream).  During CUDA-graph
    capture a fresh zeroed buffer is taken instead (its memset is captured with the launch)."""
    nbytes = max(int(nbytes), 16)
    if torch.cuda.is_current_stream_capturing():
        key = None
    else:
        key = (device.index, torch.cuda.current_stream(device).cuda_stream)
        t = _ticket_cache.get(key)
        if t is not None and t.numel() >= nbytes:
            return t
    t = torch.zeros(max(nbytes, 4096), dtype=torch.uint8, device=device)
    if key is not None:
        _ticket_cache[key] = t
    return t


def _u8(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------- layout helpers
def logits_ld(C: int) -> int:
    return _lib.load().gg_head_logits_ld(C)


def bias_pad_len(C: int) -> int:
    return _lib.load().gg_head_bias_pad(C)


# --------------------------------------------------------------------------- a1
def fuse_headings(emb: torch.Tensor, split: bool = False, want_sqnorm: bool = False):
    """(B,V,D) or (B,D) fp32 (opt-in: bf16 / fp16, read directly) -> bf16 (B,D) [or (B,3D) hi|hi|lo]; optional
    ||x||^2 (B) fp32."""
    dev = _need_cuda(emb)
    emb, in_dtype = _as_embedding(emb)
    if emb.dim() == 2:
        B, D = emb.shape
        V = 1
    else:
        B, V, D = emb.shape
    x = torch.empty((B, 3 * D if split else D), dtype=torch.bfloat16, device=emb.device)
    sq = torch.empty((B,), dtype=torch.float32, device=emb.device) if want_sqnorm else None
    lib = _lib.load()
    _call("gg_fuse_headings", lib.gg_fuse_headings, dev, _ptr(emb), in_dtype, _ptr(x), B, V, D, int(split), _ptr(sq),
          _stream(dev))
    return (x, sq) if want_sqnorm else x


_fused_cache = None  # (weakref to the embedding tensor, its version, split, x16, ||x||^2)
_FUSE_CACHE = __import__("os").environ.get("GG_FUSE_CACHE", "1") != "0"


def fuse_headings_shared(emb: torch.Tensor, split: bool = False):
    """fuse_headings(emb, split, want_sqnorm=True), remembered for the SAME tensor object at the same version: the
    serving forward of SuperGuessr and the ProtoRefiner that follows it (inference.py:168,180) both start from the
    heading mean of one embedding batch -- the second call reuses the first one's bf16 rows and norms instead of
    streaming the (B,4,D) fp32 batch from HBM again.  Returns (x16, sqnorm)."""
    global _fused_cache
    import weakref

    c = _fused_cache
    if (_FUSE_CACHE and c is not None and c[0]() is emb and c[1] == emb._version and c[2] == bool(split)
            and c[3].device == emb.device):
        return c[3], c[4]
    x16, sq = fuse_headings(emb, split=split, want_sqnorm=True)
    try:
        _fused_cache = (weakref.ref(emb), emb._version, bool(split), x16, sq) if _FUSE_CACHE else None
    except TypeError:
        _fused_cache = None
    return x16, sq


def prepare_head_weights(weight: torch.Tensor, bias: torch.Tensor, split: bool = False):
    """fp32 nn.Linear parameters -> (bf16 operand (C,D) or (C,3D), zero-padded fp32 bias)."""
    dev = _need_cuda(weight, bias)
    C, D = weight.shape
    w = weight.detach().float().contiguous()
    b = bias.detach().float().contiguous()
    w16 = torch.empty((C, 3 * D if split else D), dtype=torch.bfloat16, device=w.device)
    bp = torch.empty((bias_pad_len(C),), dtype=torch.float32, device=w.device)
    lib = _lib.load()
    _call("gg_prepare_head_weights", lib.gg_prepare_head_weights, dev, _ptr(w), _ptr(b), _ptr(w16), _ptr(bp), C, D, int(split), _stream(dev))
    return w16, bp


def fuse_and_prepare(emb: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, split: bool = False):
    """fuse_headings + prepare_head_weights of one training step in a single launch (gg_fuse_and_prepare).
    Returns (x bf16, w16 bf16, bias_pad fp32)."""
    dev = _need_cuda(emb, weight, bias)
    emb, in_dtype = _as_embedding(emb)
    if emb.dim() == 2:
        B, D = emb.shape
        V = 1
    else:
        B, V, D = emb.shape
    C = weight.shape[0]
    assert weight.shape[1] == D, "embedding dim of the batch and of the head differ"
    w = weight.detach().float().contiguous()
    b = bias.detach().float().contiguous()
    x = torch.empty((B, 3 * D if split else D), dtype=torch.bfloat16, device=emb.device)
    w16 = torch.empty((C, 3 * D if split else D), dtype=torch.bfloat16, device=emb.device)
    bp = torch.empty((bias_pad_len(C),), dtype=torch.float32, device=emb.device)
    _call("gg_fuse_and_prepare", _lib.load().gg_fuse_and_prepare, dev, _ptr(emb), in_dtype, _ptr(x), B, V, D, _ptr(w), _ptr(b),
          _ptr(w16), _ptr(bp), C, int(split), _stream(dev))
    return x, w16, bp


def row_sqnorm_bf16(m: torch.Tensor, split: bool = False) -> torch.Tensor:
    """||row||^2 of a bf16 matrix; split: rows are [hi | lo | hi] (3D entries) and stand for hi + lo."""
    dev = _need_cuda(m)
    assert m.dtype == torch.bfloat16 and m.dim() == 2 and m.is_contiguous()
    D = m.shape[1] // 3 if split else m.shape[1]
    out = torch.empty((m.shape[0],), dtype=torch.float32, device=m.device)
    if m.shape[0]:
        _call("gg_row_sqnorm_bf16", _lib.load().gg_row_sqnorm_bf16, dev, _ptr(m), m.shape[0], D, int(split), _ptr(out),
              _stream(dev))
    return out


def cast_bank_bf16(m: torch.Tensor, split: bool = False) -> torch.Tensor:
    """(rows, D) fp32 -> bf16 (rows, D), or the hi/lo split operand (rows, 3D) = [hi | lo | hi] (gg_cast_bf16)."""
    dev = _need_cuda(m)
    assert m.dtype == torch.float32 and m.dim() == 2
    m = m.contiguous()
    rows, D = m.shape
    out = torch.empty((rows, 3 * D if split else D), dtype=torch.bfloat16, device=m.device)
    if rows:
        _call("gg_cast_bf16", _lib.load().gg_cast_bf16, dev, _ptr(m), _ptr(out), rows, D, int(split), _stream(dev))
    return out


def proto_group_cells(cell_off_cpu):
    """Geocells packed into accumulation groups of <= 256 prototypes (gg_proto_group_cells, host): returns the
    (ngroups + 1) int32 group boundaries in cell indices."""
    import ctypes

    import numpy as np

    off = np.ascontiguousarray(np.asarray(cell_off_cpu, dtype=np.int32))
    ncell = len(off) - 1
    out = np.zeros(ncell + 1, dtype=np.int32)
    ng = _lib.load().gg_proto_group_cells(off.ctypes.data_as(ctypes.c_void_p), ncell, out.ctypes.data_as(ctypes.c_void_p))
    return out[: ng + 1].copy()


# --------------------------------------------------------------------------- a2-a4
def head_forward(x16, w16, bias_pad, C, k, centroids, want_logits: bool):
    """Returns dict(topk_val (B,k), topk_idx (B,k) i64, pred_cell (B) i64, pred_llh (B,2), lse (B),
    logits (B,ldc) bf16 or None)."""
    dev = _need_cuda(x16, w16, bias_pad, centroids)
    B, K = x16.shape
    assert w16.shape == (C, K), (w16.shape, C, K)
    dev = x16.device
    lib = _lib.load()
    ldc = logits_ld(C)
    logits = torch.empty((B, ldc), dtype=torch.bfloat16, device=dev) if want_logits else None
    ws = _u8(lib.gg_head_fwd_workspace_bytes(B, C, k), dev)
    tickets = _tickets(lib.gg_head_fwd_ticket_bytes(B), dev)
    topk_val = torch.empty((B, k), dtype=torch.float32, device=dev)
    topk_idx = torch.empty((B, k), dtype=torch.int64, device=dev)
    pred_cell = torch.empty((B,), dtype=torch.int64, device=dev)
    pred_llh = torch.empty((B, 2), dtype=torch.float32, device=dev)
    lse = torch.empty((B,), dtype=torch.float32, device=dev)
    _call("gg_head_fwd", lib.gg_head_fwd, dev, _ptr(x16), _ptr(w16), _ptr(bias_pad), B, C, K, _ptr(logits), ldc, k, _ptr(ws),
                        _ptr(tickets), _ptr(centroids), _ptr(topk_val), _ptr(topk_idx), _ptr(pred_cell), _ptr(pred_llh), _ptr(lse),
                        _stream(dev))
    return dict(topk_val=topk_val, topk_idx=topk_idx, pred_cell=pred_cell, pred_llh=pred_llh, lse=lse, logits=logits)


# --------------------------------------------------------------------------- a5-a8
def centroid_unit_vectors(centroids: torch.Tensor) -> torch.Tensor:
    """(C,2) (lng,lat) -> the loss kernels' centroid table (unit vectors + spatial index)."""
    dev = _need_cuda(centroids)
    c = centroids.detach().float().contiguous()
    C = c.shape[0]
    lib = _lib.load()
    table = torch.empty((lib.gg_centroid_table_floats(C),), dtype=torch.float32, device=c.device)
    ws = _u8(lib.gg_centroid_table_workspace_bytes(C), c.device)
    _call("gg_centroid_unit_vectors", lib.gg_centroid_unit_vectors, dev, _ptr(c), _ptr(table), C, _ptr(ws), _stream(dev))
    return table


def row_stats_buffer(B, C, device):
    return _u8(_lib.load().gg_hav_row_stats_bytes(B, C), device)


def hav_row_stats(labels, cent_table, C, tau=65.0, far_km=FAR_KM_DEFAULT, want_nearest=False, out=None):
    """Label-only half of the smoothed loss.  Returns (row_stats buffer, nearest_cell (B) i64 | None,
    nearest_km (B) | None).  out: a buffer from row_stats_buffer() (e.g. allocated on another stream)."""
    dev = _need_cuda(labels, cent_table)
    labels = labels.detach().float().contiguous()
    B = labels.shape[0]
    assert labels.shape == (B, 2), "labels must be (B, 2) (lng, lat)"
    dev = labels.device
    lib = _lib.load()
    stats = out if out is not None else row_stats_buffer(B, C, dev)
    ncell = torch.empty((B,), dtype=torch.int64, device=dev) if want_nearest else None
    nkm = torch.empty((B,), dtype=torch.float32, device=dev) if want_nearest else None
    _call("gg_hav_row_stats", lib.gg_hav_row_stats, dev, _ptr(labels), _ptr(cent_table), B, C, float(tau), float(far_km),
          _ptr(stats), _ptr(ncell), _ptr(nkm), _stream(dev))
    return stats, ncell, nkm


def hav_ce(logits, lse, labels, cent_xyz, C, tau=65.0, far_km=FAR_KM_DEFAULT, want_nearest=False, want_db=False,
           want_mean=False, row_stats=None):
    """Fused haversine label-smoothed CE.  Returns (dlogits bf16 (B,ldc) = p - t, loss_rows (B),
    nearest_cell (B) i64 | None, nearest_km (B) | None[, db_partials (parts, Cpad) fp32 if want_db]
    [, loss_mean () fp32 if want_mean]).  row_stats: result of hav_row_stats() when it was computed
    ahead (then labels / far_km / want_nearest are not used here)."""
    dev = _need_cuda(logits, lse, cent_xyz)
    B, ldc = logits.shape
    dev = logits.device
    lib = _lib.load()
    ncell = nkm = None
    if row_stats is None:
        row_stats, ncell, nkm = hav_row_stats(labels, cent_xyz, C, tau, far_km, want_nearest)
    dlogits = torch.empty_like(logits)
    loss_rows = torch.empty((B,), dtype=torch.float32, device=dev)
    ws = _u8(lib.gg_hav_ce_workspace_bytes(B, C), dev)
    dbp = (torch.empty((lib.gg_hav_ce_db_parts(B, C), lib.gg_hav_cpad(C)), dtype=torch.float32, device=dev)
           if want_db else None)
    mean = torch.empty((), dtype=torch.float32, device=dev) if want_mean else None
    _call("gg_hav_ce_fwd_bwd", lib.gg_hav_ce_fwd_bwd, dev, _ptr(logits), ldc, _ptr(lse), _ptr(row_stats), _ptr(cent_xyz), B, C,
          float(tau), _ptr(dlogits), _ptr(loss_rows), _ptr(dbp), _ptr(ws), _ptr(mean), 1.0 / B, _stream(dev))
    out = (dlogits, loss_rows, ncell, nkm)
    if want_db:
        out += (dbp,)
    if want_mean:
        out += (mean,)
    return out


def hard_ce(logits, lse, labels_clf, C):
    dev = _need_cuda(logits, lse, labels_clf)
    B, ldc = logits.shape
    y = labels_clf.detach().to(torch.int64).contiguous()
    assert y.shape == (B,), "labels_clf must be (B,)"
    dlogits = torch.empty_like(logits)
    loss_rows = torch.zeros((B,), dtype=torch.float32, device=logits.device)
    _call("gg_hard_ce_fwd_bwd", _lib.load().gg_hard_ce_fwd_bwd, dev, _ptr(logits), ldc, _ptr(lse), _ptr(y), B, C, _ptr(dlogits), _ptr(loss_rows),
                                       _stream(dev))
    return dlogits, loss_rows


def loss_mean(loss_rows: torch.Tensor, scale: float | None = None) -> torch.Tensor:
    dev = _need_cuda(loss_rows)
    B = loss_rows.shape[0]
    out = torch.empty((), dtype=torch.float32, device=loss_rows.device)
    _call("gg_loss_mean", _lib.load().gg_loss_mean, dev, _ptr(loss_rows), B, float(1.0 / B if scale is None else scale), _ptr(out),
                                        _stream(dev))
    return out


BWD_SCHEDULES = {None: 0, "auto": 0, "rounds": 1, "streamk": 2}  # GG_BWD_SCHEDULE_AUTO / GG_BWD_ROUNDS / GG_BWD_STREAMK


def head_backward(dlogits, x16, C, D, scale, grad_scale=None, want_db=True, db_partials=None, c_range=None, out=None,
                  push=None, schedule=None):
    """dW (C,D) fp32 = scale * grad_scale * dlogits^T x[:, :D]; db (C) (from the loss kernel's column-sum
    partials when given, else from a pass over dlogits).

    c_range=(c0, c1) computes only geocells [c0, c1) (c0 a multiple of 8) into rows c0..c1 of ``out=(dW, db)``:
    the NCCL data-parallel path runs the GEMM range by range and all-reduces each range while the next one runs.
    push=(blk_count_ptr, [ready_ptr of rank 0, 1, ...], [staging region of rank 0, 1, ...], rank): data-parallel
    mode -- every tile goes straight into its reducer's staging slab (gg_grad_exchange follows on the same stream),
    nothing is written to ``out``; whole range only.  Returns (None, None) then.
    schedule: "rounds" / "streamk" / None = auto (stream-K ranges in push mode, whole rounds + a stream-K tail else);
    the two add the same products in different orders (equal to fp32 rounding, bit-equal under the same schedule)."""
    dev = _need_cuda(dlogits, x16, grad_scale)
    B, ldc = dlogits.shape
    dev = dlogits.device
    lib = _lib.load()
    c0, c1 = (0, C) if c_range is None else c_range
    assert 0 <= c0 < c1 <= C and c0 % 8 == 0, "geocell range must start at a multiple of 8"
    dp_arr, dp_world, dp_rank = None, 0, 0
    if push is not None:
        import ctypes

        assert (c0, c1) == (0, C), "the push mode covers the whole geocell range"
        blk_count, ready, stage, dp_rank = push
        dp_world = len(ready)
        dp_arr = (ctypes.c_ulonglong * (1 + 2 * dp_world))(int(blk_count), *[int(p) for p in ready], *[int(p) for p in stage])
        dW = db = None
    elif out is None:
        dW = torch.empty((C, D), dtype=torch.float32, device=dev)
        db = torch.empty((C,), dtype=torch.float32, device=dev) if want_db else None
    else:
        dW, db = out
        if not want_db:
            db = None
    ws = _u8(lib.gg_head_bwd_workspace_bytes(c1 - c0), dev)
    if not want_db and push is None:
        db_partials = None
    if grad_scale is not None:
        grad_scale = grad_scale.detach().float().contiguous()
    esz = dlogits.element_size()
    _call("gg_head_bwd", lib.gg_head_bwd, dev, _ptr(dlogits) + c0 * esz, ldc, _ptr(x16), x16.shape[1], B, c1 - c0, D, float(scale),
          _ptr(grad_scale), 0 if dW is None else _ptr(dW) + c0 * D * 4, 0 if db is None else _ptr(db) + c0 * 4,
          0 if db_partials is None else _ptr(db_partials) + c0 * 4, 0 if db_partials is None else db_partials.shape[0],
          0 if db_partials is None else db_partials.shape[1], _ptr(ws),
          0 if dp_arr is None else ctypes.cast(dp_arr, ctypes.c_void_p), dp_world, int(dp_rank), BWD_SCHEDULES[schedule],
          _stream(dev))
    return dW, db


def head_dx(dlogits, w16, C, D, scale, grad_scale, emb_shape):
    """demb = scale * grad_scale / V * (dlogits[:, :C] @ W), broadcast over the V headings (gg_head_dx): the gradient
    with respect to the (B, V, D) / (B, D) embedding input.  w16: the forward's bf16 operand (first D columns used)."""
    dev = _need_cuda(dlogits, w16, grad_scale)
    B, ldc = dlogits.shape
    V = 1 if len(emb_shape) == 2 else int(emb_shape[1])
    demb = torch.empty(tuple(emb_shape), dtype=torch.float32, device=dlogits.device)
    if grad_scale is not None:
        grad_scale = grad_scale.detach().float().contiguous()
    _call("gg_head_dx", _lib.load().gg_head_dx, dev, _ptr(dlogits), ldc, _ptr(w16), w16.shape[1], B, C, D, float(scale),
          _ptr(grad_scale), V, _ptr(demb), _stream(dev))
    return demb


def topk_accuracy(topk_idx, targets):
    """(2,) fp32 on the device: [top-1 accuracy, top-k accuracy] of the trainer's per-step metrics
    (main_coordinator_idun_s3.py:399-408) -- no host synchronisation."""
    dev = _need_cuda(topk_idx, targets)
    topk_idx = topk_idx.detach().to(torch.int64).contiguous()
    targets = targets.detach().to(torch.int64).contiguous()
    B, k = topk_idx.shape
    assert targets.shape == (B,)
    acc = torch.empty((2,), dtype=torch.float32, device=topk_idx.device)
    _call("gg_topk_accuracy", _lib.load().gg_topk_accuracy, dev, _ptr(topk_idx), k, _ptr(targets), B, _ptr(acc), _stream(dev))
    return acc


# --------------------------------------------------------------------------- f-4 hierarchical fusion
def split3_bf16(src, role, pos_encoding=None, V=1):
    """(rows, D) fp32 -> (rows, 6D) bf16 three-term split operand (gg_split3_bf16): role 0 activation, 1 weight."""
    dev = _need_cuda(src, pos_encoding)
    assert src.dtype == torch.float32 and src.dim() == 2
    src = src.contiguous()
    rows, D = src.shape
    out = torch.empty((rows, 6 * D), dtype=torch.bfloat16, device=src.device)
    _call("gg_split3_bf16", _lib.load().gg_split3_bf16, dev, _ptr(src), rows, D, int(role), _ptr(pos_encoding), int(V),
          _ptr(out), _stream(dev))
    return out


def linear_bf16(a16, w16, bias, N):
    """out (M, N) fp32 = a16 (M, K) @ w16 (N, K)^T + bias on the tensor cores (gg_linear_bf16)."""
    dev = _need_cuda(a16, w16, bias)
    M, K = a16.shape
    assert w16.shape == (N, K) and a16.dtype == torch.bfloat16 and w16.dtype == torch.bfloat16
    out = torch.empty((M, N), dtype=torch.float32, device=a16.device)
    b = None if bias is None else bias.detach().float().contiguous()
    _call("gg_linear_bf16", _lib.load().gg_linear_bf16, dev, _ptr(a16), K, _ptr(w16), K, _ptr(b), M, N, K, _ptr(out), N,
          _stream(dev))
    return out


def hier_fuse(x, in_proj_weight, in_proj_bias, out_proj_weight, out_proj_bias, pos_encoding, num_heads=16, weights=None):
    """`hierarchical=True` fusion in eval mode (super_guessr.py:340-345): x (B, V, D) fp32 -> (B, D) fp32.
    weights: optional cached (in_proj split, out_proj split) operands (they only change with the parameters)."""
    dev = _need_cuda(x, in_proj_weight, out_proj_weight, pos_encoding)
    assert x.dtype == torch.float32 and x.dim() == 3
    B, V, D = x.shape
    pe = pos_encoding.detach().float().reshape(-1, D).contiguous()
    if B > pe.shape[0]:
        raise RuntimeError(f"hierarchical fusion: the positional table is indexed by the batch row and holds "
                           f"{pe.shape[0]} rows (models/layers/positional_encoder.py:44); batch {B} does not broadcast")
    if weights is None:
        weights = (split3_bf16(in_proj_weight.detach().float(), 1), split3_bf16(out_proj_weight.detach().float(), 1))
    w_in, w_out = weights
    zs = split3_bf16(x.contiguous().view(B * V, D), 0, pos_encoding=pe, V=V)
    qkv = linear_bf16(zs, w_in, in_proj_bias, 3 * D)
    ctx = torch.empty((B, 6 * D), dtype=torch.bfloat16, device=x.device)
    _call("gg_hier_attention", _lib.load().gg_hier_attention, dev, _ptr(qkv), B, V, D, int(num_heads), _ptr(ctx), _stream(dev))
    return linear_bf16(ctx, w_out, out_proj_bias, D)


GRAD_CTRL_BYTES = 16384  # GG_GRAD_CTRL_BYTES
GRAD_CTRL_READY_OFF = 4096  # GG_GRAD_CTRL_READY_OFF


def grad_stage_floats(C, D, world):
    lib = _lib.load()
    assert lib.gg_grad_ctrl_bytes() == GRAD_CTRL_BYTES, "control-region size of the library and of the binding differ"
    return int(lib.gg_grad_stage_floats(int(C), int(D), int(world)))


def grad_exchange(grad_ptrs, ctrl_ptrs, grad_mc, ctrl_mc, stage_ptr, rank, C, D, no_wait=False):
    """Second half of the fused gradient exchange (gg_grad_exchange), on the current stream, after the dW GEMM that
    pushed the tiles: per owned block add the staged copies (local memory) in rank order and write the average into
    every rank's gradient buffer.  grad_ptrs / ctrl_ptrs: every rank's buffer / control region as mapped on this
    device; grad_mc / ctrl_mc: their multicast addresses (0 = posted peer stores); stage_ptr: this rank's staging
    region.  no_wait: GG_GRAD_NO_WAIT (emulation)."""
    import ctypes

    world = len(grad_ptrs)
    g = (ctypes.c_ulonglong * world)(*[int(p) for p in grad_ptrs])
    c = (ctypes.c_ulonglong * world)(*[int(p) for p in ctrl_ptrs])
    dev = None  # raw addresses: the caller's current device / stream
    _call("gg_grad_exchange", _lib.load().gg_grad_exchange, dev, ctypes.cast(g, ctypes.c_void_p),
          ctypes.cast(c, ctypes.c_void_p), int(grad_mc), int(ctrl_mc), int(stage_ptr), world, int(rank), int(C), int(D),
          int(bool(no_wait)), _stream(dev))


def grad_exchange_adamw(w16_ptrs, bias_ptrs, ctrl_ptrs, w16_mc, bias_mc, ctrl_mc, stage_ptr, rank, C, D, master_w,
                        master_b, m_w, v_w, m_b, v_b, hyper, step, no_wait=False):
    """The fused exchange with the optimizer in it (gg_grad_exchange_adamw): per owned block add the staged copies,
    apply AdamW to this rank's rows of the fp32 master weights / bias (in place, with the moments) and write the bf16
    operand rows and the fp32 bias into every rank's operand buffers.  hyper: device fp32 [lr, beta1, beta2, eps,
    weight_decay]; step: device int64 counter (incremented by the kernel)."""
    import ctypes

    world = len(w16_ptrs)
    dev = _need_cuda(master_w, master_b, m_w, v_w, m_b, v_b, hyper, step)
    for t in (master_w, master_b, m_w, v_w, m_b, v_b, hyper):
        assert t.dtype == torch.float32 and t.is_contiguous()
    assert step.dtype == torch.int64 and hyper.numel() >= 5
    assert master_w.shape == (C, D) and m_w.shape == (C, D) and v_w.shape == (C, D)
    w = (ctypes.c_ulonglong * world)(*[int(p) for p in w16_ptrs])
    b = (ctypes.c_ulonglong * world)(*[int(p) for p in bias_ptrs])
    c = (ctypes.c_ulonglong * world)(*[int(p) for p in ctrl_ptrs])
    _call("gg_grad_exchange_adamw", _lib.load().gg_grad_exchange_adamw, dev, ctypes.cast(w, ctypes.c_void_p),
          ctypes.cast(b, ctypes.c_void_p), ctypes.cast(c, ctypes.c_void_p), int(w16_mc), int(bias_mc), int(ctrl_mc),
          int(stage_ptr), world, int(rank), int(C), int(D), _ptr(master_w), _ptr(master_b), _ptr(m_w), _ptr(v_w),
          _ptr(m_b), _ptr(v_b), _ptr(hyper), _ptr(step), int(bool(no_wait)), _stream(dev))


def p2p_allreduce_avg(peer_ptrs, rank, n_floats):
    """In-place two-shot average of a symmetric-memory fp32 buffer over the ranks (gg_p2p_allreduce_avg) on the
    current stream.  peer_ptrs: every rank's buffer address as mapped on this device (rank order).  The caller
    puts a symmetric-memory barrier on either side."""
    import ctypes

    world = len(peer_ptrs)
    arr = (ctypes.c_ulonglong * world)(*[int(p) for p in peer_ptrs])
    dev = None  # raw addresses: the caller's current device / stream
    _call("gg_p2p_allreduce_avg", _lib.load().gg_p2p_allreduce_avg, dev, ctypes.cast(arr, ctypes.c_void_p), world, int(rank),
          int(n_floats), _stream(dev))


def nvls_allreduce_avg(multicast_ptr, world, rank, n_floats):
    """The same average through the NVSwitch multicast mapping of the buffer (gg_nvls_allreduce_avg)."""
    dev = None  # raw address: the caller's current device / stream
    _call("gg_nvls_allreduce_avg", _lib.load().gg_nvls_allreduce_avg, dev, int(multicast_ptr), int(world), int(rank),
          int(n_floats), _stream(dev))


def p2p_slice(n_floats, world, rank):
    """[lo, hi) of rank's slice of the gradient buffer, in floats (gg_p2p_slice)."""
    import ctypes

    lo, hi = ctypes.c_size_t(0), ctypes.c_size_t(0)
    _lib.load().gg_p2p_slice(int(n_floats), int(world), int(rank), ctypes.byref(lo), ctypes.byref(hi))
    return 4 * lo.value, 4 * hi.value


# --------------------------------------------------------------------------- a10-a15
METRICS = {"l2": 0, "cosine": 1}  # GG_METRIC_L2 / GG_METRIC_COSINE


def proto_retrieve(q16, q_sqnorm, cand, topk, bank16, bank_sqnorm, bank_coords, cell_off, cell_lo, cell_hi,
                   proto_base=0, group_off=None, metric="l2", gather4=True, want_meta=False):
    """Stage 0+1.  Returns the (B*topk, 4) fp32 record array {score, lng, lat, proto id bits} [and the launcher's
    meta words (int32 x 4, device): work items, pairs, accumulation units].  group_off: device int32 group
    boundaries (proto_group_cells); None = one group per cell."""
    dev = _need_cuda(q16, q_sqnorm, cand, cell_off, bank16, group_off)
    B, D = q16.shape
    lib = _lib.load()
    cand = cand.detach().to(torch.int64).contiguous()
    n_protos = 0 if bank16 is None else bank16.shape[0]
    if group_off is None:
        group_off = torch.arange(cell_hi - cell_lo + 1, dtype=torch.int32, device=dev)
    rec = torch.empty((B * topk, 4), dtype=torch.float32, device=dev)
    ws = _u8(lib.gg_proto_retrieve_workspace_bytes(B, topk, D, cell_hi - cell_lo), dev)
    _call("gg_proto_retrieve", lib.gg_proto_retrieve, dev, _ptr(q16), _ptr(q_sqnorm), B, D, _ptr(cand), cand.shape[1], topk,
          _ptr(bank16), _ptr(bank_sqnorm), _ptr(bank_coords), n_protos, _ptr(cell_off), cell_lo, cell_hi, _ptr(group_off),
          group_off.numel() - 1, proto_base, METRICS[metric], 0 if gather4 else 1, _ptr(rec), _ptr(ws), _stream(dev))
    if want_meta:
        return rec, ws[:16].view(torch.int32)
    return rec


def proto_record_ids(rec):
    """int64 prototype ids of a (n, 4) record array (gg_proto_record_ids): the candidate list of the image stage."""
    dev = _need_cuda(rec)
    n = rec.shape[0]
    ids = torch.empty((n,), dtype=torch.int64, device=rec.device)
    _call("gg_proto_record_ids", _lib.load().gg_proto_record_ids, dev, _ptr(rec), n, _ptr(ids), _stream(dev))
    return ids


def proto_take_image_coords(rec, rec_img):
    """rec[i].{lng, lat} = rec_img[i].{lng, lat} where the image stage found a member (gg_proto_take_image_coords)."""
    dev = _need_cuda(rec, rec_img)
    assert rec.shape == rec_img.shape
    _call("gg_proto_take_image_coords", _lib.load().gg_proto_take_image_coords, dev, _ptr(rec), _ptr(rec_img), rec.shape[0],
          _stream(dev))
    return rec


def proto_refine(rec, nranks, cand, cand_probs, initial, topk, temperature, max_refinement, want_debug=False):
    """Stage 2.  rec: (nranks, B*topk, 4) or (B*topk, 4).  Returns (preds_LLH (B,2) f32, preds_geocell (B) i64,
    guess_index (B) i32[, score (B,topk), proto (B,topk) i32])."""
    dev = _need_cuda(rec, cand, initial)
    dev = rec.device
    cand = cand.detach().to(torch.int64).contiguous()
    B = cand.shape[0]
    initial = initial.detach().float().contiguous()
    assert initial.shape == (B, 2)
    if cand_probs is not None:
        cand_probs = cand_probs.detach().float().contiguous()
    rec = rec.contiguous()
    out_llh = torch.empty((B, 2), dtype=torch.float32, device=dev)
    out_cell = torch.empty((B,), dtype=torch.int64, device=dev)
    out_guess = torch.empty((B,), dtype=torch.int32, device=dev)
    out_score = torch.empty((B, topk), dtype=torch.float32, device=dev) if want_debug else None
    out_proto = torch.empty((B, topk), dtype=torch.int32, device=dev) if want_debug else None
    _call("gg_proto_refine", _lib.load().gg_proto_refine, dev, _ptr(rec), nranks, B * topk, _ptr(cand_probs),
                                    0 if cand_probs is None else cand_probs.shape[1], _ptr(cand), cand.shape[1],
                                    _ptr(initial), B, topk, float(temperature), float(max_refinement), _ptr(out_llh),
                                    _ptr(out_cell), _ptr(out_guess), _ptr(out_score), _ptr(out_proto), _stream(dev))
    if want_debug:
        return out_llh, out_cell, out_guess, out_score, out_proto
    return out_llh, out_cell, out_guess
