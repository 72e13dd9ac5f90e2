"""AdamW for the geocell head, sharded over the data-parallel ranks and fused into the gradient exchange.

The reference trainer runs ``torch.optim.AdamW(model.parameters(), lr, betas, weight_decay)`` after DDP's all-reduce
(main_coordinator_idun_s3.py:286-291,423-424; training/train_eval_loop.py:188-202,241).  Both halves of that -- the
all-gather half of the all-reduce and the optimizer pass -- move the whole fp32 head (51.8 MB at D = 1024) on every
rank every step.  Here the rank that REDUCES a block of 128 geocells also OWNS its optimizer state:

    loss.backward()   gg_head_bwd pushes every dW / db tile into the owner's staging slab (NVLink, under the GEMM);
                      gg_grad_exchange_adamw adds the staged copies in rank order, applies AdamW to the owner's rows
                      of the fp32 master weights and writes the new rows AS THE bf16 OPERAND the next forward reads
                      (and the fp32 bias) into every rank's operand buffer -- half the bytes of the averaged gradient;
    optimizer.step()  bookkeeping only: the update already happened.

What changes for the caller: ``cell_layer.weight.grad`` stays ``None`` (no gradient is materialised); the fp32
``cell_layer.weight`` / ``.bias`` of a rank are current only for the blocks it owns until ``gather_master()`` is called
(do that on ALL ranks before ``state_dict()`` / checkpointing); one ``backward()`` per ``step()`` (no gradient
accumulation).  Learning-rate schedulers work on ``param_groups`` as usual: the hyper-parameters are re-read at every
``backward()``.  The arithmetic is torch's fused AdamW kernel in fp32 (tests/test_dp_gpu.py compares a step with
NCCL all-reduce + torch.optim.AdamW).
"""
from __future__ import annotations

import torch

from . import ops

_BLOCK = 128  # geocells per exchanged block (gg_grad_exchange: block b belongs to rank b % world)


class ShardedAdamW(torch.optim.Optimizer):
    def __init__(self, module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, process_group=None):
        import torch.distributed as dist

        w, b = module.cell_layer.weight, module.cell_layer.bias
        if not w.is_cuda:
            raise RuntimeError("ShardedAdamW runs on the CUDA path only (move the model to the GPU first)")
        if module.precision != "bf16":
            raise ValueError("ShardedAdamW keeps the plain bf16 operand; precision='bf16x3' is not covered")
        if not (w.requires_grad and b.requires_grad):
            raise ValueError("ShardedAdamW updates cell_layer.weight and .bias: both must require grad")
        super().__init__([w, b], dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.module = module
        dev = w.device
        C, D = w.shape
        if D % 8:
            raise ValueError("embedding dim must be a multiple of 8")
        world, rank, group = 1, 0, None
        if dist.is_initialized():
            group = process_group if process_group is not None else (
                module._dp["group"] if module._dp is not None and module._dp["group"] is not None else dist.group.WORLD)
            world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world not in (1, 2, 4, 8):
            raise ValueError(f"ShardedAdamW covers 1, 2, 4 or 8 ranks of one NVSwitch domain; this group has {world}")
        self.world, self.rank, self.group, self.C, self.D = world, rank, group, C, D

        ctrl_words = ops.GRAD_CTRL_BYTES // 4
        n_stage = ops.grad_stage_floats(C, D, world)
        n_w16 = C * D // 2                       # bf16 (C, D) counted in floats; D % 8 == 0 keeps 16-byte alignment
        n_w16 = -(-n_w16 // 4) * 4
        n_bias = -(-ops.bias_pad_len(C) // 4) * 4
        total = ctrl_words + n_stage + n_w16 + n_bias
        mc = 0
        if world > 1:
            import torch.distributed._symmetric_memory as symm

            dist.broadcast(w.data, dist.get_global_rank(group, 0), group=group)  # one starting point on every rank
            dist.broadcast(b.data, dist.get_global_rank(group, 0), group=group)
            buf = symm.empty(total, dtype=torch.float32, device=dev)
            handle = symm.rendezvous(buf, group.group_name)
            base = [int(p) for p in handle.buffer_ptrs]
            mc = int(getattr(handle, "multicast_ptr", 0) or 0)
        else:
            buf = torch.empty(total, dtype=torch.float32, device=dev)
            handle = None
            base = [buf.data_ptr()]
        buf.zero_()
        o_stage, o_w16, o_bias = ctrl_words, ctrl_words + n_stage, ctrl_words + n_stage + n_w16
        self.buf, self.handle = buf, handle
        self.w16 = buf[o_w16: o_w16 + C * D // 2].view(torch.bfloat16).view(C, D)
        self.bias_pad = buf[o_bias: o_bias + ops.bias_pad_len(C)]
        w16_0, bias_0 = ops.prepare_head_weights(w, b)
        self.w16.copy_(w16_0)
        self.bias_pad.copy_(bias_0)
        torch.cuda.current_stream(dev).synchronize()
        if handle is not None:
            handle.barrier(channel=0)  # every rank's control words are zero and its operands in place
        self.ctrl_ptrs = base
        self.stage_ptrs = [p + 4 * o_stage for p in base]
        self.w16_ptrs = [p + 4 * o_w16 for p in base]
        self.bias_ptrs = [p + 4 * o_bias for p in base]
        self.use_mc = mc != 0 and world >= 4  # as the plain fused exchange: multicast stores pay from 4 ranks on
        self.mc = (mc, mc + 4 * o_w16, mc + 4 * o_bias) if self.use_mc else (0, 0, 0)
        self.m_w, self.v_w = torch.zeros_like(w.data), torch.zeros_like(w.data)
        self.m_b, self.v_b = torch.zeros_like(b.data), torch.zeros_like(b.data)
        self.hyper = torch.zeros(8, dtype=torch.float32, device=dev)
        self.step_count = torch.zeros((), dtype=torch.int64, device=dev)
        self._hyper_host = None
        self._pending = 0
        self._sync_hyper()
        module._sharded = self
        module._op_cache = None

    # ---- what the module calls -------------------------------------------------------------------------------
    def operands(self):
        return dict(w16=self.w16, bias_pad=self.bias_pad, split=False)

    def _sync_hyper(self):
        g = self.param_groups[0]
        vals = (float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]))
        if vals != self._hyper_host:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("ShardedAdamW: hyper-parameters changed inside a CUDA-graph capture; change them "
                                   "between replays (the graph reads them from device memory)")
            self.hyper[:5].copy_(torch.tensor(vals, dtype=torch.float32))
            self._hyper_host = vals

    def backward_and_step(self, dlogits, x16, C, D, B, gloss, dbp):
        """Called from the head's autograd backward: dW GEMM with the push epilogue, then exchange + AdamW."""
        if self._pending:
            raise RuntimeError("ShardedAdamW applies the update inside backward(): call optimizer.step() after every "
                               "backward() (gradient accumulation is not covered)")
        self._sync_hyper()
        w, b = self.module.cell_layer.weight, self.module.cell_layer.bias
        ready = [p + ops.GRAD_CTRL_READY_OFF for p in self.ctrl_ptrs]
        ops.head_backward(dlogits, x16, C, D, scale=1.0 / B, grad_scale=gloss, want_db=True, db_partials=dbp,
                          push=(self.ctrl_ptrs[self.rank], ready, self.stage_ptrs, self.rank))
        ops.grad_exchange_adamw(self.w16_ptrs, self.bias_ptrs, self.ctrl_ptrs, self.mc[1], self.mc[2], self.mc[0],
                                self.stage_ptrs[self.rank], self.rank, C, D, w.data, b.data, self.m_w, self.v_w,
                                self.m_b, self.v_b, self.hyper, self.step_count)
        if not torch.cuda.is_current_stream_capturing():
            self._pending = 1

    # ---- torch.optim.Optimizer surface -----------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._pending = 0
        return loss

    def zero_grad(self, set_to_none: bool = True):
        for p in (self.module.cell_layer.weight, self.module.cell_layer.bias):
            p.grad = None

    @staticmethod
    def owned_mask(C: int, world: int, rank: int, device=None):
        """Boolean (C,) mask of the geocell rows rank `rank` of `world` reduces and updates: blocks of 128 geocells,
        block b belongs to rank b % world (the rule gg_head_bwd's push mode and gg_grad_exchange_adamw share)."""
        rows = torch.arange(C, device=device)
        return (rows // _BLOCK) % world == rank

    def owned_rows(self):
        """Boolean (C,) mask of the geocell rows whose master weights / moments this rank keeps current."""
        return self.owned_mask(self.C, self.world, self.rank, self.hyper.device)

    @torch.no_grad()
    def gather_master(self):
        """Make the fp32 ``cell_layer.weight`` / ``.bias`` current on EVERY rank (collective; call it on all ranks
        before state_dict() / evaluation code that reads the fp32 parameters).  The moments stay sharded."""
        if self.world == 1:
            return
        import torch.distributed as dist

        mask = self.owned_rows()
        for p in (self.module.cell_layer.weight, self.module.cell_layer.bias):
            m = mask.view(-1, *([1] * (p.dim() - 1)))
            t = torch.where(m, p.data, torch.zeros((), dtype=p.dtype, device=p.device))
            dist.all_reduce(t, group=self.group)
            p.data.copy_(t)

    def state_dict(self):
        """param_groups as torch's optimizers + this rank's shard of the moments (rows of owned_rows()) and the step."""
        sd = super().state_dict()
        mask = self.owned_rows()
        sd["sharded"] = dict(step=int(self.step_count.item()), rank=self.rank, world=self.world,
                             exp_avg_w=self.m_w[mask].clone(), exp_avg_sq_w=self.v_w[mask].clone(),
                             exp_avg_b=self.m_b[mask].clone(), exp_avg_sq_b=self.v_b[mask].clone())
        return sd

    def load_state_dict(self, sd):
        sh = sd.get("sharded")
        super().load_state_dict({k: v for k, v in sd.items() if k != "sharded"})
        if sh is not None:
            if (sh["rank"], sh["world"]) != (self.rank, self.world):
                raise ValueError("ShardedAdamW state belongs to another rank / world size")
            mask = self.owned_rows()
            self.m_w[mask] = sh["exp_avg_w"].to(self.m_w.device)
            self.v_w[mask] = sh["exp_avg_sq_w"].to(self.v_w.device)
            self.m_b[mask] = sh["exp_avg_b"].to(self.m_b.device)
            self.v_b[mask] = sh["exp_avg_sq_b"].to(self.v_b.device)
            self.step_count.fill_(int(sh["step"]))
        self._hyper_host = None
        self._sync_hyper()

    def describe(self):
        how = "multimem.st through the NVSwitch" if self.use_mc else ("posted peer stores" if self.world > 1 else "local stores")
        return (f"sharded AdamW over {self.world} rank(s): gg_head_bwd pushes every dW / db tile into the owning rank's "
                "staging slab (fp32, under the GEMM); gg_grad_exchange_adamw adds the staged copies in rank order, "
                "updates the owner's fp32 master rows and moments and broadcasts the bf16 operand rows + fp32 bias "
                f"({how}); no gradient all-gather, no per-rank optimizer pass, no per-step weight recast")
