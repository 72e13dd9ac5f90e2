"""SuperGuessr: drop-in for the reference's geocell classifier (models/super_guessr.py:20-395)
whose post-encoder path runs on hand-written sm_100a kernels.

Kept from the reference: constructor keywords, ``forward`` signature and return types
(``ModelOutput`` in training / eval, the 3-tuple when ``serving`` in eval mode), the attribute
names callers read (``geocell_centroid_coords``, ``num_cells``, ``cell_layer``, ``serving``,
``num_candidates``) and the state-dict keys (``cell_layer.weight`` (C,D), ``cell_layer.bias`` (C),
``geocell_centroid_coords`` (C,2)), so checkpoints filtered by name and shape (inference.py:134-156)
load unchanged and torch optimisers / wandb.watch see ordinary ``.grad`` tensors.

What runs where:
  a1 fusion            -> gg_fuse_headings          a5-a7 smoothed CE      -> gg_hav_ce_fwd_bwd
  a2-a4 head + top-k   -> gg_head_fwd (tcgen05)     a8    dW, db           -> gg_head_bwd (tcgen05)
The vision encoder (base_model) stays PyTorch and is outside this path.
"""
from __future__ import annotations

import os

import torch
from torch import Tensor, nn

from . import ops
from .geocells import resolve_centroids
from .utils import CLIP_EMBED_DIM, CLIP_PRETRAINED_HEAD, LABEL_SMOOTHING_CONSTANT, ModelOutput, TopK


class _GeocellHeadLoss(torch.autograd.Function):
    """fusion -> head GEMM (+top-k) -> fused CE forward+gradient; backward = the dW/db GEMM."""

    @staticmethod
    def forward(ctx, embedding, weight, bias, module, labels, labels_clf, smooth):
        C, D = weight.shape
        # The label-only half of the loss (nearest centroid, sum of smoothed targets, near-group mask) reads no
        # logits and runs on a side stream.  Launched AFTER the persistent head GEMM (its small CTAs co-reside
        # with the GEMM's one CTA per SM and fill its idle issue slots) but ordered only behind the work that
        # precedes the GEMM, so the HBM-bound fusion / weight cast have the machine to themselves.
        stats = None
        serial = ops.timing_enabled()  # per-launcher timing (bench.py): nothing runs beside the kernel being timed
        under_gemm = smooth and module._stats_under_gemm and not serial
        if smooth and serial:
            stats = module._row_stats_prepare(labels, C)[3]
            ops.hav_row_stats(labels.detach().float().contiguous(), module._centroid_xyz(), C,
                              tau=module.label_smoothing_tau, far_km=module.far_km, out=stats)
        elif smooth and not under_gemm:
            stats = module._row_stats_async(labels, C)
        if module._sharded is not None:
            # sharded AdamW: the bf16 operand is persistent (the reducers of the last step wrote it): fusion only
            st = module._sharded.operands()
            x16 = ops.fuse_headings(embedding)
        elif module.training:
            # the weights move every step: their bf16 operand is rebuilt in the same launch as the fusion
            split = module.precision == "bf16x3"
            x16, w16, bias_pad = ops.fuse_and_prepare(embedding, weight, bias, split=split)
            module._op_cache = None  # the eval-mode cache must not outlive a training step
            st = dict(w16=w16, bias_pad=bias_pad, split=split)
        else:
            st = module._operands(weight, bias)
            x16 = ops.fuse_headings(embedding, split=st["split"])
        if under_gemm:
            fork = module._row_stats_prepare(labels, C)
        head = ops.head_forward(x16, st["w16"], st["bias_pad"], C, module.num_candidates,
                                module.geocell_centroid_coords.data, want_logits=True)
        if under_gemm:
            stats = module._row_stats_launch(fork)
        dbp = None
        if smooth:
            torch.cuda.current_stream().wait_stream(module._side_stream)
            dlogits, loss_rows, _, _, dbp, loss = ops.hav_ce(head["logits"], head["lse"], None, module._centroid_xyz(),
                                                             C, tau=module.label_smoothing_tau, want_db=True,
                                                             want_mean=True, row_stats=stats)
        else:
            dlogits, loss_rows = ops.hard_ce(head["logits"], head["lse"], labels_clf, C)
            loss = ops.loss_mean(loss_rows)
        ctx.save_for_backward(dlogits, x16)
        ctx.w16 = st["w16"] if embedding.requires_grad else None  # dx = dlogits W reuses the forward's operand
        ctx.dbp = dbp
        ctx.set_materialize_grads(False)  # no zero-filled grads for the non-differentiable outputs
        ctx.dims = (C, D, embedding.shape, embedding.requires_grad)
        ctx.weight = weight
        ctx.module = module
        outs = (head["topk_val"], head["topk_idx"], head["pred_cell"], head["pred_llh"])
        ctx.mark_non_differentiable(*outs)
        return (loss,) + outs

    @staticmethod
    def backward(ctx, gloss, *_):
        dlogits, x16 = ctx.saved_tensors
        C, D, emb_shape, emb_needs_grad = ctx.dims
        B = dlogits.shape[0]
        want_w, want_b = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        dW = db = demb = None
        sharded = ctx.module._sharded is not None and ctx.module.training
        if sharded and ctx.needs_input_grad[0] and emb_needs_grad:
            # the reducers overwrite this rank's operand as soon as it has announced a block: dx reads W first
            demb = ops.head_dx(dlogits, ctx.w16, C, D, 1.0 / B, gloss, emb_shape)
        if sharded:
            if not (want_w and want_b):
                raise RuntimeError("sharded AdamW updates cell_layer.weight and .bias together: both must require grad")
            ctx.module._sharded.backward_and_step(dlogits, x16, C, D, B, gloss, ctx.dbp)
            return demb, None, None, None, None, None, None
        if want_w or want_b:
            dp = ctx.module._dp
            if dp is None:
                dW, db = ops.head_backward(dlogits, x16, C, D, scale=1.0 / B, grad_scale=gloss, want_db=want_b,
                                           db_partials=ctx.dbp)
            else:
                dW, db = ctx.module._backward_data_parallel(dlogits, x16, C, D, B, gloss, want_b, ctx.dbp)
        if ctx.needs_input_grad[0] and emb_needs_grad:
            # Reached when the encoder is trained (TinyViT's last stage / CLIP's last layer, super_guessr.py:127-153;
            # main_coordinator_idun_s3.py:423): dx = dlogits W on the tensor cores with the mean's 1/V broadcast
            # (super_guessr.py:347) in the epilogue (gg_head_dx).
            demb = ops.head_dx(dlogits, ctx.w16, C, D, 1.0 / B, gloss, emb_shape)
        return demb, (dW if want_w else None), db, None, None, None, None


NUM_ATTENTION_HEADS = 16  # models/super_guessr.py:14


class PositionalEncoder(nn.Module):
    """The reference's positional table (models/layers/positional_encoder.py:5-44): (max_len, 1, D), sin on even
    columns, cos on odd ones, registered as a frozen parameter `pos_encoding` (same state-dict key).  The addition
    itself -- `x + pos_encoding[:B]`, which indexes the BATCH row -- runs inside gg_split3_bf16."""

    def __init__(self, dim_model: int, dropout_p: float = 0.1, max_len: int = 1000):
        super().__init__()
        import math

        self.dropout = nn.Dropout(dropout_p)
        pos_encoding = torch.zeros(max_len, dim_model)
        positions = torch.arange(0, max_len, dtype=torch.float).view(-1, 1)
        division = torch.exp(torch.arange(0, dim_model, 2).float() * (-math.log(10000.0)) / dim_model)
        pos_encoding[:, 0::2] = torch.sin(positions * division)
        pos_encoding[:, 1::2] = torch.cos(positions * division)
        self.register_parameter("pos_encoding", nn.Parameter(pos_encoding.unsqueeze(0).transpose(0, 1), requires_grad=False))


def dp_chunk_bounds(C: int, chunks: int, align: int = 256):
    """Geocell ranges [c0, c1) for the chunked dW GEMM + all-reduce: boundaries on multiples of ``align``
    (one CTA pair's 2 x 128 geocell blocks), sizes as equal as the alignment allows, never empty."""
    blocks = -(-C // align)
    chunks = max(1, min(int(chunks), blocks))
    base, rem = divmod(blocks, chunks)
    out, b0 = [], 0
    for i in range(chunks):
        b1 = b0 + base + (1 if i < rem else 0)
        out.append((b0 * align, min(b1 * align, C)))
        b0 = b1
    return out


def _all_reduce_avg(t: Tensor, group, comm_dtype):
    import torch.distributed as dist

    c = t if comm_dtype is None or comm_dtype == t.dtype else t.to(comm_dtype)
    if c.is_cuda:
        dist.all_reduce(c, op=dist.ReduceOp.AVG, group=group)  # NCCL
    else:  # gloo (CPU tests of the host logic) has no AVG
        dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
        c /= dist.get_world_size(group)
    if c is not t:
        t.copy_(c)


class SuperGuessr(nn.Module):
    def __init__(
        self,
        base_model: nn.Module = None,
        panorama: bool = False,
        hierarchical: bool = False,
        should_smooth_labels: bool = False,
        serving: bool = False,
        freeze_base: bool = False,
        num_candidates: int = 5,
        embed_dim: int = CLIP_EMBED_DIM,
        *,
        centroids=None,
        precision: str = "bf16",
        label_smoothing_tau: float = LABEL_SMOOTHING_CONSTANT,
        far_km: float = ops.FAR_KM_DEFAULT,
        **kwargs,
    ):
        """Same arguments as the reference (super_guessr.py:21-54).  Keyword-only additions:

        centroids: explicit (C,2) (lng,lat) table instead of the one resolved from
            ``data/geocells`` / the packaged copy (geocells.resolve_centroids).
        precision: "bf16" (operands rounded to bf16, fp32 accumulate; BASELINE cfg2/cfg4) or
            "bf16x3" (hi/lo split operands, ~fp32-faithful logits for fp32 checkpoints; cfg1).
        label_smoothing_tau: config.LABEL_SMOOTHING_CONSTANT (65 km).
        far_km: target-mass cut-off of the fused loss (see gg_hav_ce_fwd_bwd); inf = exact.
        """
        super().__init__()
        if len(kwargs) > 0:  # reference behaviour, super_guessr.py:57-59
            print(f"Not using keyword arguments: {list(kwargs.keys())}")
        if precision not in ("bf16", "bf16x3"):
            raise ValueError("precision must be 'bf16' or 'bf16x3'")
        self.base_model = base_model
        self.panorama = panorama
        self.hidden_size = embed_dim
        self.serving = serving
        self.should_smooth_labels = should_smooth_labels
        self.freeze_base = freeze_base
        self.hierarchical = hierarchical
        self.num_candidates = num_candidates
        self.precision = precision
        self.label_smoothing_tau = float(label_smoothing_tau)
        self.far_km = float(far_km)
        self._set_hidden_size()

        table = resolve_centroids(centroids)
        self.geocell_centroid_coords = nn.Parameter(table, requires_grad=False)
        self.num_cells = table.size(0)
        self.input_dim = self.hidden_size
        if self.hierarchical:  # super_guessr.py:88-99 (same sub-modules, hence the same state-dict keys)
            print("Number of attention heads:", NUM_ATTENTION_HEADS)
            self.heading_pad = 0
            self.pos_encoder = PositionalEncoder(self.input_dim + self.heading_pad)
            self.self_attn = nn.MultiheadAttention(self.input_dim + self.heading_pad, NUM_ATTENTION_HEADS, dropout=0.1,
                                                   batch_first=True)
            self.relu = nn.ReLU()
        self._hier_cache = None
        self.cell_layer = nn.Linear(self.input_dim, self.num_cells)
        self.softmax = nn.Softmax(dim=-1)
        self.loss_fnc = nn.CrossEntropyLoss()
        self._freeze_params()
        self._op_cache = None
        self._xyz_cache = None
        self._side_stream = None
        self._stats_under_gemm = os.environ.get("GG_STATS_UNDER_GEMM", "1") != "0"
        self._dp = None
        self._sharded = None
        print(f"Initialized SuperGuessr classification model with {self.num_cells} geocells.")

    # ---- reference helpers (super_guessr.py:114-206) ---------------------------------------
    def _set_hidden_size(self):
        if self.base_model is not None:
            try:
                self.hidden_size = self.base_model.config.hidden_size
                self.mode = "transformer"
            except AttributeError:
                self.hidden_size = self.base_model.config.hidden_sizes[-1]
                self.mode = "convnext"

    def _freeze_params(self):
        """Same three branches as the reference (super_guessr.py:127-153): ``freeze_base`` freezes the whole
        encoder; a CLIP backbone in training loads the pretrained head (if the file exists) and freezes every
        encoder layer but the last; a TinyViT backbone in training freezes all but its last stage."""
        if self.base_model is None:
            return
        if self.freeze_base:
            for p in self.base_model.parameters():
                p.requires_grad = False
            return
        name = str(getattr(getattr(self.base_model, "config", None), "_name_or_path", ""))
        if "clip-vit" in name and not self.serving:
            head = CLIP_PRETRAINED_HEAD
            if os.path.exists(head):
                self.load_state(head)
                print(f"Initialized model parameters from model: {head}")
                for p in self.base_model.vision_model.encoder.layers[:-1].parameters():
                    p.requires_grad = False
            else:
                print(f"Warning: pretrained head not found at '{head}'. "
                      "Proceeding without loading and without freezing base layers.")
        elif "tiny" in name and not self.serving:
            self.base_model.freeze_all_but_last_stage()

    def _move_to_cuda(self, pixel_values=None, embedding=None, labels=None, labels_clf=None):
        # the reference only moves inputs in eval mode (:187-199)
        if not self.training and next(self.parameters()).is_cuda:
            dev = next(self.parameters()).device
            pixel_values, embedding, labels, labels_clf = (
                t.to(dev) if t is not None else None for t in (pixel_values, embedding, labels, labels_clf))
        return pixel_values, embedding, labels, labels_clf

    def load_state(self, path: str):
        own = self.state_dict()
        for name, param in torch.load(path, map_location="cuda" if torch.cuda.is_available() else "cpu").items():
            if name not in own:
                print(f"Parameter {name} not in model's state.")
                continue
            own[name].copy_(param.data if isinstance(param, nn.Parameter) else param)

    # ---- operand caches ---------------------------------------------------------------------
    def _operands(self, weight, bias):
        """bf16 (or hi/lo split) copy of the head weights + padded bias.  In training mode it is
        rebuilt every step (fused optimisers update parameters without bumping ``_version``, so a
        version key cannot be trusted while weights are moving); in eval mode it is cached and
        rebuilt when the fp32 parameters change (load_state_dict, .to(), manual edits)."""
        if self._sharded is not None:
            return self._sharded.operands()
        key = (weight.data_ptr(), weight._version, bias.data_ptr(), bias._version, self.precision, weight.device)
        if self.training or self._op_cache is None or self._op_cache["key"] != key:
            split = self.precision == "bf16x3"
            w16, bias_pad = ops.prepare_head_weights(weight, bias, split=split)
            self._op_cache = dict(key=key, w16=w16, bias_pad=bias_pad, split=split)
        return self._op_cache

    def _row_stats_prepare(self, labels, C):
        """Buffers for gg_hav_row_stats, made on the current stream, and the fork point (an event on the
        current stream) the side-stream launch is ordered behind."""
        cur = torch.cuda.current_stream()
        table = self._centroid_xyz()
        if self._side_stream is None or self._side_stream.device != labels.device:
            self._side_stream = torch.cuda.Stream(device=labels.device)
        # the buffer belongs to the main stream (which joins the side stream before the loss kernel reads
        # it and before anything later could reuse it); labels are made fp32-contiguous there as well
        labels = labels.detach().float().contiguous()
        stats = ops.row_stats_buffer(labels.shape[0], C, labels.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        return labels, table, C, stats, ev

    def _row_stats_launch(self, fork):
        """gg_hav_row_stats on the side stream behind the fork point (joined by the caller before the loss kernel)."""
        labels, table, C, stats, ev = fork
        side = self._side_stream
        side.wait_event(ev)
        with torch.cuda.stream(side):
            ops.hav_row_stats(labels, table, C, tau=self.label_smoothing_tau, far_km=self.far_km, out=stats)
        return stats

    def _row_stats_async(self, labels, C):
        return self._row_stats_launch(self._row_stats_prepare(labels, C))

    def _centroid_xyz(self):
        c = self.geocell_centroid_coords
        key = (c.data_ptr(), c._version, c.device)
        if self._xyz_cache is None or self._xyz_cache[0] != key:
            self._xyz_cache = (key, ops.centroid_unit_vectors(c.data))
        return self._xyz_cache[1]

    # ---- data-parallel head training --------------------------------------------
    def enable_data_parallel(self, process_group=None, chunks: int = 1, comm_dtype=None, comm: str = "auto"):
        """Average the head gradients over the ranks of ``process_group`` INSIDE ``backward()`` (what the
        reference leaves to DDP / Accelerate).  After this call ``.grad`` is already the global average: do not
        wrap the module in DDP as well.

        comm="fused": [dW | db] and a staging region live in a symmetric-memory buffer mapped into every peer over
        NVLink.  The dW GEMM stores every finished tile straight into the staging slab of the rank that reduces its
        block of 128 geocells (posted stores: the transfer rides under the GEMM) and announces complete blocks;
        gg_grad_exchange, behind it on the same stream, adds the locally staged copies in rank order and writes the
        averages into every rank's gradient (NVSwitch multicast stores from 8 ranks, posted peer stores below).
        No peer loads, no host barrier.
        comm="nvls" / "p2p": the same buffer, ONE exchange kernel after the GEMM between two symmetric-memory
        barriers -- "nvls" through the NVSwitch multicast mapping (gg_nvls_allreduce_avg), "p2p" with peer loads /
        stores (gg_p2p_allreduce_avg: two-shot, rank-order sums).
        comm="nccl": NCCL all-reduce (AVG) of dW and db on a communication stream; with ``chunks`` > 1 the dW GEMM
        runs in that many geocell ranges (dp_chunk_bounds), each all-reduced while the next one is computed.
        ``comm_dtype=torch.bfloat16`` (NCCL only) halves the bytes on NVLink by rounding each rank's gradient
        before the sum (opt-in: not bit-faithful to an fp32 all-reduce).
        comm="auto": what measured fastest on B200 NVSwitch boxes for the 51.9 MB head gradient -- "fused" at 8
        ranks, "p2p" at 2 and 4 -- when symmetric memory can be set up, else "nccl" (also for CPU / gloo groups,
        where the path uses whatever all_reduce the group's backend provides).  The transport is decided when the
        buffer is set up, before any kernel runs."""
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("enable_data_parallel needs an initialised torch.distributed process group")
        if comm not in ("auto", "fused", "nvls", "p2p", "nccl"):
            raise ValueError("comm must be 'auto', 'fused', 'nvls', 'p2p' or 'nccl'")
        self._dp = dict(group=process_group, chunks=int(chunks), comm_dtype=comm_dtype, stream=None, comm=comm,
                        symm=None)
        return self

    def sharded_adamw(self, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, process_group=None):
        """The optimizer for the head in place of ``torch.optim.AdamW(model.parameters(), ...)``
        (main_coordinator_idun_s3.py:286-291): AdamW sharded over the data-parallel ranks and fused into the gradient
        exchange (sharded_adamw.ShardedAdamW).  With an initialised process group the head is data parallel over it
        (do not also call enable_data_parallel / wrap in DDP); without one it is a single-GPU fused optimizer step.
        Returns a torch.optim.Optimizer: ``loss.backward(); optimizer.step()`` as before."""
        from .sharded_adamw import ShardedAdamW

        return ShardedAdamW(self, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, process_group=process_group)

    @staticmethod
    def _dp_transport(comm: str, world: int, is_cuda: bool, comm_dtype=None) -> str:
        """Which gradient exchange a group gets (pure function; host-tested): own kernels only cover 2, 4 or 8
        ranks of one NVSwitch domain, anything else is NCCL -- decided before the first kernel runs."""
        if comm == "nccl" or not is_cuda or comm_dtype is not None:
            return "nccl"
        if world in (2, 4, 8):
            # measured on B200 NVSwitch boxes (profiles/README.md, r02): at 8 ranks the exchange fused into the dW
            # GEMM wins (0.470 ms per step against 0.515 multicast / 0.522 peer two-shot / 0.593 NCCL); at 2 and 4
            # ranks a rank's NVLink egress carries its gradient twice in the fused scheme (push + broadcast) and the
            # two-shot exchange after the GEMM, which uses both directions at once, is as fast or faster
            if comm == "auto":
                return "fused" if world >= 8 else "p2p"
            return comm
        if comm == "auto":
            return "nccl"
        raise ValueError(f"comm='{comm}' covers 2, 4 or 8 ranks; this group has {world} (use comm='auto' or 'nccl')")

    def _symm_gradient_buffer(self, C, D, dev):
        """The symmetric-memory [control | dW | db | pad] buffer and its peer addresses (allocated, zeroed and
        exchanged once).  Returns None when the group gets NCCL instead."""
        import torch.distributed as dist

        dp = self._dp
        if dp["symm"] is not None and dp["symm"]["key"] == (C, D, dev):
            return dp["symm"]
        if dp["comm"] == "nccl":
            return None
        group = dp["group"] if dp["group"] is not None else dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        kind = self._dp_transport(dp["comm"], world, dev.type == "cuda", dp["comm_dtype"])
        if kind == "nccl":
            dp["comm"] = "nccl"
            return None
        try:
            import torch.distributed._symmetric_memory as symm

            n = C * D + C
            n_pad = -(-n // (4 * world)) * (4 * world)
            ctrl_words = ops.GRAD_CTRL_BYTES // 4
            n_stage = ops.grad_stage_floats(C, D, world) if kind == "fused" else 0
            buf = symm.empty(ctrl_words + n_pad + n_stage, dtype=torch.float32, device=dev)
            handle = symm.rendezvous(buf, group.group_name)
        except Exception as e:  # noqa: BLE001
            if dp["comm"] != "auto":
                raise
            import sys

            print(f"[geoguessr_ai_b200] symmetric-memory gradient exchange unavailable ({type(e).__name__}: {e}); "
                  "using NCCL all-reduce", file=sys.stderr)
            dp["comm"] = "nccl"
            return None
        mc = int(getattr(handle, "multicast_ptr", 0) or 0)
        if kind == "nvls" and mc == 0:
            raise RuntimeError("comm='nvls': this group has no NVSwitch multicast mapping (multicast_ptr == 0)")
        buf.zero_()  # control words, pad and staging rows start at zero on every rank ...
        torch.cuda.current_stream(dev).synchronize()
        handle.barrier(channel=0)  # ... before any peer can add to them
        base = [int(p) for p in handle.buffer_ptrs]
        off = ops.GRAD_CTRL_BYTES
        use_mc = mc != 0 and (kind == "nvls" or (kind == "fused" and world >= 4))
        dp["symm"] = dict(key=(C, D, dev), buf=buf, handle=handle, ctrl_ptrs=base, ptrs=[p + off for p in base],
                          stage_ptrs=[p + off + 4 * n_pad for p in base],
                          world=world, rank=rank, n=n_pad, multicast=(mc + off) if mc else 0,
                          ctrl_multicast=mc, kind=kind, use_mc=use_mc, grad_off=ctrl_words)
        return dp["symm"]

    def describe_data_parallel(self):
        """One line for logs / the bench record: which exchange runs."""
        dp = self._dp
        if dp is None:
            return None
        sm = dp.get("symm")
        if sm is None:
            return (f"nccl avg, {'bf16' if dp['comm_dtype'] is not None else 'fp32'}, {dp['chunks']} geocell range(s) "
                    "overlapped with the dW GEMM")
        path = ("NVSwitch multicast (multimem.ld_reduce / multimem.st)" if sm["use_mc"]
                else "peer loads in rank order + peer stores")
        if sm["kind"] == "fused":
            how = "multimem.st through the NVSwitch" if sm["use_mc"] else "posted peer stores"
            return ("fused: gg_head_bwd stores every dW / db tile into the reducing rank's staging slab (symmetric memory, "
                    "fp32, posted over NVLink under the GEMM) and announces complete 128-geocell blocks; gg_grad_exchange "
                    f"adds the staged copies in rank order and broadcasts the averages ({how}); no peer loads, no host barrier")
        return (f"{sm['kind']}: one exchange kernel ({path}) over [dW | db] in symmetric memory after the dW GEMM, "
                "between two symmetric-memory barriers")

    def _grad_views(self, sm, C, D):
        g0 = sm["grad_off"]
        buf = sm["buf"]
        dW = buf[g0: g0 + C * D].view(C, D)
        db = buf[g0 + C * D: g0 + C * D + C]
        # gradient accumulation without zero_grad(): .grad may still alias this buffer -- move it out first
        for prm, view in ((self.cell_layer.weight, dW), (self.cell_layer.bias, db)):
            if prm.grad is not None and prm.grad.data_ptr() == view.data_ptr():
                prm.grad = prm.grad.clone()
        return dW, db

    def _backward_data_parallel_symm(self, sm, dlogits, x16, C, D, B, gloss, want_b, dbp):
        dW, db = self._grad_views(sm, C, D)
        dp = self._dp
        dev = dlogits.device
        if dp["stream"] is None or dp["stream"].device != dev:
            dp["stream"] = torch.cuda.Stream(device=dev)
        comm, cur = dp["stream"], torch.cuda.current_stream(dev)
        if sm["kind"] == "fused":
            # the dW GEMM pushes every tile into its reducer's staging slab (over NVLink, under the GEMM) and announces
            # complete blocks; the exchange kernel behind it adds the staged copies and broadcasts the averages
            ready = [p + ops.GRAD_CTRL_READY_OFF for p in sm["ctrl_ptrs"]]
            ops.head_backward(dlogits, x16, C, D, scale=1.0 / B, grad_scale=gloss, want_db=True, db_partials=dbp,
                              push=(sm["ctrl_ptrs"][sm["rank"]], ready, sm["stage_ptrs"], sm["rank"]))
            ops.grad_exchange(sm["ptrs"], sm["ctrl_ptrs"], sm["multicast"] if sm["use_mc"] else 0,
                              sm["ctrl_multicast"] if sm["use_mc"] else 0, sm["stage_ptrs"][sm["rank"]], sm["rank"], C, D)
            return dW, (db if want_b else None)

        h = sm["handle"]
        ops.head_backward(dlogits, x16, C, D, scale=1.0 / B, grad_scale=gloss, want_db=True, db_partials=dbp,
                          out=(dW, db))
        comm.wait_stream(cur)
        with torch.cuda.stream(comm):
            h.barrier(channel=0)  # every rank's gradient is written
            if sm["use_mc"]:
                ops.nvls_allreduce_avg(sm["multicast"], sm["world"], sm["rank"], sm["n"])
            else:
                ops.p2p_allreduce_avg(sm["ptrs"], sm["rank"], sm["n"])
            h.barrier(channel=1)  # every rank's slice has landed everywhere (and nobody still reads my copy)
        cur.wait_stream(comm)
        return dW, (db if want_b else None)

    def _backward_data_parallel(self, dlogits, x16, C, D, B, gloss, want_b, dbp):
        dp = self._dp
        dev = dlogits.device
        sm = self._symm_gradient_buffer(C, D, dev) if dlogits.is_cuda else None
        if sm is not None:
            return self._backward_data_parallel_symm(sm, dlogits, x16, C, D, B, gloss, want_b, dbp)
        if dp["stream"] is None or dp["stream"].device != dev:
            dp["stream"] = torch.cuda.Stream(device=dev)
        comm, cur = dp["stream"], torch.cuda.current_stream()
        dW = torch.empty((C, D), dtype=torch.float32, device=dev)
        db = torch.empty((C,), dtype=torch.float32, device=dev) if want_b else None
        comm.wait_stream(cur)  # orders the allocator's reuse of dW / db memory behind earlier work
        for c0, c1 in dp_chunk_bounds(C, dp["chunks"]):
            ops.head_backward(dlogits, x16, C, D, scale=1.0 / B, grad_scale=gloss, want_db=want_b, db_partials=dbp,
                              c_range=(c0, c1), out=(dW, db))
            done = torch.cuda.Event()
            done.record(cur)
            comm.wait_event(done)
            with torch.cuda.stream(comm):
                _all_reduce_avg(dW[c0:c1], dp["group"], dp["comm_dtype"])
        if db is not None:
            with torch.cuda.stream(comm):
                _all_reduce_avg(db, dp["group"], None)
        cur.wait_stream(comm)
        return dW, db

    # ---- hierarchical fusion (super_guessr.py:340-345) and trainer-side helpers ---------------
    def _hierarchical_fusion(self, x):
        """pos_encoder + 16-head self-attention over the headings, token 0, in eval mode (gg_split3_bf16,
        gg_linear_bf16, gg_hier_attention).  Training this branch (its two dropouts, the attention backward) is not
        on the accelerated path: no reference caller enables `hierarchical`."""
        if self.training:
            raise NotImplementedError(
                "hierarchical=True is implemented for eval mode (dropout off, attention parameters frozen); training "
                "the attention fusion is outside the accelerated path (no reference caller enables it)")
        a = self.self_attn
        key = tuple((p.data_ptr(), p._version) for p in (a.in_proj_weight, a.out_proj.weight))
        if self._hier_cache is None or self._hier_cache[0] != key:
            self._hier_cache = (key, (ops.split3_bf16(a.in_proj_weight.detach().float(), 1),
                                      ops.split3_bf16(a.out_proj.weight.detach().float(), 1)))
        with torch.no_grad():
            return ops.hier_fuse(x, a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias,
                                 self.pos_encoder.pos_encoding, num_heads=a.num_heads, weights=self._hier_cache[1])

    def labels_from_coords(self, labels: Tensor):
        """The trainer's label derivation (main_coordinator_idun_s3.py:390-391: haversine_matrix + argmin) as a
        by-product of the loss's row statistics: (labels_clf (B,) int64, nearest_km (B,) fp32), first index on ties."""
        labels = labels.to(self.geocell_centroid_coords.device)
        _, cell, km = ops.hav_row_stats(labels, self._centroid_xyz(), self.num_cells, tau=self.label_smoothing_tau,
                                        far_km=self.far_km, want_nearest=True)
        return cell, km

    @staticmethod
    def accuracy(topk, targets: Tensor) -> Tensor:
        """The trainer's per-step metrics (main_coordinator_idun_s3.py:399-408) on the device: a (2,) tensor
        [top-1 accuracy, top-k accuracy]; read it when logging instead of two `.item()` syncs per step."""
        idx = topk.indices if hasattr(topk, "indices") else topk
        return ops.topk_accuracy(idx, targets.to(idx.device))

    # ---- forward (super_guessr.py:268-395) ---------------------------------------------------
    def forward(self, pixel_values: Tensor = None, embedding: Tensor = None, labels: Tensor = None,
                labels_clf: Tensor = None, index: Tensor = None):
        if self.base_model is not None:
            assert pixel_values is not None, 'Parameter "pixel_values" must be supplied if model has a base model.'
        else:
            assert embedding is not None, 'Parameter "embedding" must be supplied if model does not have a base model.'
        pixel_values, embedding, labels, labels_clf = self._move_to_cuda(pixel_values, embedding, labels, labels_clf)

        # encoder (outside the accelerated path; same handling as :308-334)
        if self.base_model is not None and pixel_values is not None:
            if self.panorama:
                assert pixel_values.dim() == 5, "panorama=True expects (B, 4, C, H, W)"
                n, v, c, h, w = pixel_values.shape
                pixel_values = pixel_values.view(n * v, c, h, w)
            elif pixel_values.dim() > 4:
                pixel_values = pixel_values.squeeze(1)
            outs = self.base_model(pixel_values=pixel_values)
            if hasattr(outs, "last_hidden_state") and self.mode == "transformer":
                embedding = outs.last_hidden_state.mean(dim=1)
            elif hasattr(outs, "pooler_output"):
                embedding = outs.pooler_output
            else:
                embedding = outs
            if self.panorama:
                embedding = embedding.view(n, v, -1)

        layer_input = embedding
        if self.panorama:
            assert layer_input.dim() == 3, "panorama=True expects embeddings of shape (B, headings, D)"
        else:
            assert layer_input.dim() == 2, "panorama=False expects embeddings of shape (B, D)"
        weight, bias = self.cell_layer.weight, self.cell_layer.bias
        if not weight.is_cuda:
            raise ops._lib.GeoguessrB200Error(
                "SuperGuessr parameters are on the CPU: this implementation only runs on an sm_100a GPU "
                "(call .cuda()); there is no CPU fallback")
        if self.panorama and self.hierarchical:
            layer_input = self._hierarchical_fusion(layer_input)  # (B, D) fp32: the head sees a single "heading"

        serving_now = (not self.training) and self.serving
        needs_loss = not serving_now
        smooth = bool(getattr(self, "should_smooth_labels", False)) and labels is not None
        if needs_loss and not smooth and labels_clf is None:
            raise ValueError("labels_clf is required for the hard-label loss (should_smooth_labels=False or labels=None)")

        if needs_loss:
            # the fused gradient costs one extra bf16 write even under no_grad; still one code path
            loss, tv, ti, pred_cell, pred_llh = _GeocellHeadLoss.apply(
                layer_input, weight, bias, self, labels, labels_clf, smooth)
            topk = TopK(tv, ti)
            return ModelOutput(loss, loss, pred_llh, pred_cell, topk, embedding)

        with torch.no_grad():
            st = self._operands(weight, bias)
            # (shared with a ProtoRefiner called on the same embedding batch next: one pass over the fp32 batch)
            x16, _ = ops.fuse_headings_shared(layer_input, split=st["split"])
            head = ops.head_forward(x16, st["w16"], st["bias_pad"], self.num_cells, self.num_candidates,
                                    self.geocell_centroid_coords.data, want_logits=False)
        return head["pred_llh"], TopK(head["topk_val"], head["topk_idx"]), embedding

    def __str__(self):
        rep = "SuperGuessr(\n"
        rep += f"\tbase_model\t= {self.base_model is not None}\n"
        rep += f"\tpanorama\t= {self.panorama}\n"
        rep += f"\thierarchical\t= {self.hierarchical}\n"
        rep += f"\tembedding_size\t= {self.hidden_size}\n"
        rep += f"\tinput_dim\t= {self.input_dim}\n"
        rep += f"\tnum_geocells\t= {self.num_cells}\n"
        rep += f"\tlabel_smoothing\t= {self.should_smooth_labels}\n"
        rep += f"\tfreeze_base\t= {self.freeze_base}\n"
        rep += f"\tserving\t\t= {self.serving}\n"
        rep += ")"
        return rep
