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
