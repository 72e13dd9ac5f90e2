#!/usr/bin/env python
"""Probe (torchrun, N GPUs): does symmetric memory + NVLS multicast work on this box, and what do a 52 MB fp32
all-reduce cost through NCCL and through torch's multimem op?"""
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 12647 * 1024


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


g = torch.randn(n, device=dev)
t_nccl = timeit(lambda: dist.all_reduce(g, op=dist.ReduceOp.AVG))
g16 = g.bfloat16()
t_nccl16 = timeit(lambda: dist.all_reduce(g16, op=dist.ReduceOp.AVG))
if rank == 0:
    print(f"nccl all_reduce fp32 {n * 4 / 1e6:.1f} MB: {t_nccl:.1f} us; bf16: {t_nccl16:.1f} us", flush=True)
try:
    t = symm.empty(n, dtype=torch.float32, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD.group_name)
    if rank == 0:
        print("symm rendezvous ok; multicast_ptr", hex(h.multicast_ptr), "buffer_ptrs", [hex(p) for p in h.buffer_ptrs],
              "signal_pad_size", h.signal_pad_size, "has_multicast", symm._SymmetricMemory.has_multicast_support(
                  torch._C._autograd.DeviceType.CUDA if hasattr(torch._C._autograd, 'DeviceType') else 1, local)
              if False else "n/a", flush=True)
    t.copy_(g)
    name = dist.group.WORLD.group_name
    t_mm = timeit(lambda: torch.ops.symm_mem.multimem_all_reduce_(t, "sum", name))
    t_two = timeit(lambda: torch.ops.symm_mem.two_shot_all_reduce_(t, "sum", name))
    tb = timeit(lambda: h.barrier(channel=0))
    if rank == 0:
        print(f"multimem_all_reduce_: {t_mm:.1f} us; two_shot: {t_two:.1f} us; symm barrier: {tb:.1f} us", flush=True)
    # what one direction of NVLink gives for half the buffer (the slice a two-rank exchange moves each way):
    # the copy engine (cudaMemcpyPeer behind Tensor.copy_) against the same bytes stored by SM threads
    peer = (rank + 1) % world
    half = n // 2
    remote = h.get_buffer(peer, (n,), torch.float32)
    src = torch.randn(half, device=dev)
    t_dma = timeit(lambda: remote[:half].copy_(src))
    t_dma_in = timeit(lambda: src.copy_(remote[:half]))
    dist.barrier()
    if rank == 0:
        print(f"{half * 4 / 1e6:.1f} MB to the peer: copy engine push {t_dma:.1f} us = {half * 4 / t_dma / 1e3:.0f} GB/s, "
              f"pull {t_dma_in:.1f} us = {half * 4 / t_dma_in / 1e3:.0f} GB/s", flush=True)
    # the same bytes stored by the SMs, by the form of the store (gg_debug_nvlink_store_probe): every rank stores into
    # its neighbour at once, as the fused gradient exchange does
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from geoguessr_ai_b200 import _lib  # noqa: E402

    lib = _lib.load()
    nbytes = half * 4 // (1 << 16) * (1 << 16)
    dst = int(h.buffer_ptrs[peer])
    stream = torch.cuda.current_stream().cuda_stream
    for mode, chunk, ctas in ((0, 0, 148), (0, 0, 592), (1, 4096, 148), (1, 32768, 148), (2, 0, 148), (2, 0, 296)):
        def store():
            rc = lib.gg_debug_nvlink_store_probe(dst, nbytes, mode, chunk, ctas, stream)
            assert rc == 0, lib.gg_last_error().decode()
        t_s = timeit(store)
        dist.barrier()
        if rank == 0:
            what = ("coalesced 16-byte st.global" if mode == 0 else f"cp.async.bulk, {chunk} B per instruction" if mode == 1
                    else "cp.async.bulk.tensor 2-D, 32 x 32 fp32 boxes of a (rows, 1024) matrix")
            print(f"{nbytes / 1e6:.1f} MB to the peer by SM stores ({what}, {ctas} CTAs): {t_s:.1f} us = "
                  f"{nbytes / t_s / 1e3:.0f} GB/s", flush=True)
    local = torch.empty(n, dtype=torch.float32, device=dev)
    for mode, chunk, ctas in ((0, 0, 592), (1, 32768, 148), (2, 0, 148)):
        def store_local():
            lib.gg_debug_nvlink_store_probe(local.data_ptr(), nbytes, mode, chunk, ctas, stream)
        t_s = timeit(store_local)
        if rank == 0:
            print(f"  (same kernel into local HBM, mode {mode}: {t_s:.1f} us = {nbytes / t_s / 1e3:.0f} GB/s)", flush=True)
except Exception as e:  # noqa: BLE001
    if rank == 0:
        print("symmetric memory probe failed:", type(e).__name__, str(e)[:500], flush=True)
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
