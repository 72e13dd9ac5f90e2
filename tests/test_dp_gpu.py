"""Data-parallel gradient exchange: the two-shot peer-memory all-reduce kernel.

On one GPU the kernel is exercised with `world` buffers of the same device standing in for the peers' copies
(peer pointers are plain addresses): launching every rank's slice in turn must leave the rank-order average in
every copy, bit for bit.  With two or more GPUs a torchrun job checks the real symmetric-memory path against an
NCCL all-reduce and a data-parallel SuperGuessr step against the unsharded one."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,n", [(2, 12647 * 64 + 12648), (4, 4096), (8, 8 * 4 * 37 + 16), (2, 8), (8, 32)])
def test_two_shot_allreduce_single_device_emulation(world, n):
    from geoguessr_ai_b200 import ops

    assert n % 4 == 0
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(world * 1000 + n)
    bufs = [torch.randn(n, generator=g).to(dev) for _ in range(world)]
    want = bufs[0].clone()
    for r in range(1, world):
        want += bufs[r]  # rank order, fp32
    want *= 1.0 / world
    ptrs = [b.data_ptr() for b in bufs]
    covered = torch.zeros(n, dtype=torch.int32)
    for r in range(world):
        lo, hi = ops.p2p_slice(n, world, r)
        covered[lo:hi] += 1
        ops.p2p_allreduce_avg(ptrs, r, n)
    torch.cuda.synchronize()
    assert torch.all(covered == 1), "slices must cover every element exactly once"
    for r in range(world):
        assert torch.equal(bufs[r], want), f"copy {r} differs from the rank-order average"


@pytest.mark.parametrize("world,Cc,D,B", [(2, 1500, 192, 200), (4, 1500, 576, 200), (8, 2100, 256, 136), (2, 12647, 1024, 512)])
def test_fused_exchange_single_device_emulation(world, Cc, D, B):
    """gg_head_bwd pushing its tiles into the reducers' staging slabs + gg_grad_exchange adding the staged copies,
    with `world` buffers (control | gradient | staging) of ONE GPU standing in for the ranks (peer pointers are plain
    addresses).  Two steps back to back (the counters only grow: epochs), with / without the loss kernel's db
    partials.  Every copy must end up holding the rank-order average of the per-rank gradients, bit for bit."""
    from geoguessr_ai_b200 import ops

    dev = torch.device("cuda:0")
    ctrl_words = ops.GRAD_CTRL_BYTES // 4
    n = Cc * D + Cc
    n_pad = -(-n // (4 * world)) * (4 * world)
    n_stage = ops.grad_stage_floats(Cc, D, world)
    bufs = [torch.zeros(ctrl_words + n_pad + n_stage, dtype=torch.float32, device=dev) for _ in range(world)]
    ctrl_ptrs = [b.data_ptr() for b in bufs]
    grad_ptrs = [p + ops.GRAD_CTRL_BYTES for p in ctrl_ptrs]
    stage_ptrs = [p + 4 * n_pad for p in grad_ptrs]
    ready = [p + ops.GRAD_CTRL_READY_OFF for p in ctrl_ptrs]
    ldc = ops.logits_ld(Cc)
    for step in range(2):
        g = torch.Generator().manual_seed(100 * world + step)
        xs, dls, dbps, want_w, want_b = [], [], [], None, None
        for r in range(world):
            x = (torch.randn(B, D, generator=g) * 0.5).to(torch.bfloat16).to(dev)
            dl = torch.zeros((B, ldc), dtype=torch.bfloat16)
            dl[:, :Cc] = (torch.randn(B, Cc, generator=g) * 1e-2).to(torch.bfloat16)
            dl = dl.to(dev)
            dbp = torch.randn(3, ldc, generator=g).to(dev) if step == 0 else None  # with / without the loss partials
            xs.append(x)
            dls.append(dl)
            dbps.append(dbp)
            dW, db = ops.head_backward(dl, x, Cc, D, scale=0.5, db_partials=dbp)
            want_w = dW if want_w is None else want_w + dW  # rank order, fp32
            want_b = db if want_b is None else want_b + db
        want_w, want_b = want_w * (1.0 / world), want_b * (1.0 / world)
        for b in bufs:
            b[ctrl_words: ctrl_words + n_pad].fill_(float("nan"))  # the gradient must be fully overwritten
        torch.cuda.synchronize()
        # one stream: every "rank" pushes, then every "rank" reduces what it owns (GG_GRAD_NO_WAIT: on ONE GPU the
        # exchange kernels run one after the other and must not wait for each other's blocks)
        for r in range(world):
            ops.head_backward(dls[r], xs[r], Cc, D, scale=0.5, db_partials=dbps[r],
                              push=(ctrl_ptrs[r], ready, stage_ptrs, r))
        for r in range(world):
            ops.grad_exchange(grad_ptrs, ctrl_ptrs, 0, 0, stage_ptrs[r], r, Cc, D, no_wait=True)
        torch.cuda.synchronize()
        for r in range(world):
            dW = bufs[r][ctrl_words: ctrl_words + Cc * D].view(Cc, D)
            db = bufs[r][ctrl_words + Cc * D: ctrl_words + Cc * D + Cc]
            assert torch.equal(dW, want_w), f"step {step}: dW of copy {r} differs from the rank-order average"
            assert torch.equal(db, want_b), f"step {step}: db of copy {r} differs"


def test_allreduce_argument_errors():
    from geoguessr_ai_b200 import ops
    from geoguessr_ai_b200._lib import GeoguessrB200Error

    b = torch.zeros(16, device="cuda:0")
    with pytest.raises(GeoguessrB200Error):
        ops.p2p_allreduce_avg([b.data_ptr()] * 3, 0, 16)  # 3 ranks: unsupported
    with pytest.raises(GeoguessrB200Error):
        ops.p2p_allreduce_avg([b.data_ptr(), 0], 0, 16)  # missing peer
    with pytest.raises(GeoguessrB200Error):
        ops.p2p_allreduce_avg([b.data_ptr(), b.data_ptr()], 0, 10)  # not a multiple of 4
    ops.p2p_allreduce_avg([b.data_ptr()], 0, 16)  # one rank: nothing to do


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_symmetric_memory_exchange_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(REPO, "tools", "p2p_check.py"), "--quick"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    assert r.returncode == 0 and "p2p_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
