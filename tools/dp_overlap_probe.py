#!/usr/bin/env python
"""Single-GPU probe of the fused gradient exchange's cost structure at cfg2 (B=4096, D=1024, C=12647):
 (a) gg_head_bwd alone, (b) two emulated ranks on this GPU: both GEMMs in push mode (tiles into the staging slabs)
 followed by both exchange kernels (sum of the staged copies, averages into both gradient buffers).
Times are CUDA events on the GEMM's stream (us)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geoguessr_ai_b200 import ops  # noqa: E402

B, D, C = 4096, 1024, 12647
dev = torch.device("cuda:0")
ldc = ops.logits_ld(C)
g = torch.Generator(device=dev).manual_seed(0)
x = (torch.randn((B, D), device=dev, generator=g) * 0.5).to(torch.bfloat16)
dl = (torch.randn((B, ldc), device=dev, generator=g) * 1e-2).to(torch.bfloat16)
world = 2
ctrl_words = ops.GRAD_CTRL_BYTES // 4
n = C * D + C
n_pad = -(-n // (4 * world)) * (4 * world)
n_stage = ops.grad_stage_floats(C, D, world)
bufs = [torch.zeros(ctrl_words + n_pad + n_stage, dtype=torch.float32, device=dev) for _ in range(world)]
ctrl = [b.data_ptr() for b in bufs]
grad = [p + ops.GRAD_CTRL_BYTES for p in ctrl]
stage = [p + 4 * n_pad for p in grad]
ready = [p + ops.GRAD_CTRL_READY_OFF for p in ctrl]
views = [(b[ctrl_words: ctrl_words + C * D].view(C, D), b[ctrl_words + C * D: ctrl_words + C * D + C]) for b in bufs]


def timed(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        torch.cuda.synchronize()
        torch.cuda._sleep(400_000)  # the host enqueues everything while the device is still busy
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def plain():
    ops.head_backward(dl, x, C, D, 1.0 / B, out=views[0])


def push_pair():  # both "ranks" push, nobody reduces
    for r in range(world):
        ops.head_backward(dl, x, C, D, 1.0 / B, push=(ctrl[r], ready, stage, r))


def reduce_pair():
    for r in range(world):
        ops.grad_exchange(grad, ctrl, 0, 0, stage[r], r, C, D, no_wait=True)


t_plain = timed(plain)
print(f"gg_head_bwd alone (rounds + stream-K tail)      {t_plain:8.1f} us")
t_sk = timed(lambda: ops.head_backward(dl, x, C, D, 1.0 / B, out=views[0], schedule="streamk"))
print(f"gg_head_bwd alone (stream-K ranges)             {t_sk:8.1f} us")


def push_one():  # rank 0's GEMM in push mode; every slab is local memory here: the cost of the push code itself
    ops.head_backward(dl, x, C, D, 1.0 / B, push=(ctrl[0], ready, stage, 0))
    for b in bufs:  # (the counters only make sense with both ranks pushing: reset what one push leaves behind)
        b[:ctrl_words].zero_()


t_push1 = timed(push_one)
print(f"gg_head_bwd in push mode, slabs local (+ reset)  {t_push1:8.1f} us")


def both():
    push_pair()
    reduce_pair()


t_both = timed(both, iters=6)
print(f"2 x gg_head_bwd (push) + 2 x gg_grad_exchange    {t_both:8.1f} us  (two emulated ranks on one GPU: per rank ~ half)")
