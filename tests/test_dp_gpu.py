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
            dW, db = ops.head_backward(dl, x, Cc, D, scale=0.5, db_partials=dbp, schedule="streamk")  # the push mode's split
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


def _torch_adamw(w, b, gw, gb, steps_done, lr, betas, eps, wd, state):
    """One torch.optim.AdamW step (the reference trainer's optimizer) on clones; state carries the moments."""
    if "opt" not in state:
        state["w"] = torch.nn.Parameter(w.clone())
        state["b"] = torch.nn.Parameter(b.clone())
        state["opt"] = torch.optim.AdamW([state["w"], state["b"]], lr=lr, betas=betas, eps=eps, weight_decay=wd)
    state["w"].grad, state["b"].grad = gw.clone(), gb.clone()
    state["opt"].step()
    return state["w"].data, state["b"].data


@pytest.mark.parametrize("world,Cc,D,B", [(1, 700, 64, 96), (2, 1500, 192, 200), (4, 1500, 576, 136), (8, 2100, 256, 136)])
def test_exchange_with_sharded_adamw_single_device_emulation(world, Cc, D, B):
    """gg_head_bwd (push) + gg_grad_exchange_adamw with `world` buffer sets of ONE GPU standing in for the ranks: after
    each of three steps every copy of the bf16 operand / fp32 bias must hold what NCCL-average + torch.optim.AdamW
    gives (master weights to fp32 rounding of a different operation order, operand = their bf16 rounding), and each
    rank's master rows are current exactly for the blocks it owns."""
    from geoguessr_ai_b200 import ops

    dev = torch.device("cuda:0")
    lr, betas, eps, wd = 3e-3, (0.9, 0.95), 1e-8, 0.05
    ctrl_words = ops.GRAD_CTRL_BYTES // 4
    n_stage = ops.grad_stage_floats(Cc, D, world)
    n_w16 = -(-(Cc * D // 2) // 4) * 4
    Cpad = ops.bias_pad_len(Cc)
    n_bias = -(-Cpad // 4) * 4
    bufs = [torch.zeros(ctrl_words + n_stage + n_w16 + n_bias, dtype=torch.float32, device=dev) for _ in range(world)]
    ctrl_ptrs = [b.data_ptr() for b in bufs]
    stage_ptrs = [p + 4 * ctrl_words for p in ctrl_ptrs]
    w16_ptrs = [p + 4 * (ctrl_words + n_stage) for p in ctrl_ptrs]
    bias_ptrs = [p + 4 * (ctrl_words + n_stage + n_w16) for p in ctrl_ptrs]
    ready = [p + ops.GRAD_CTRL_READY_OFF for p in ctrl_ptrs]
    g = torch.Generator().manual_seed(7 + world)
    w0 = (torch.randn(Cc, D, generator=g) * 0.05).to(dev)
    b0 = (torch.randn(Cc, generator=g) * 0.05).to(dev)
    masters = [(w0.clone(), b0.clone()) for _ in range(world)]
    moments = [tuple(torch.zeros_like(t) for t in (w0, w0, b0, b0)) for _ in range(world)]
    hyper = torch.tensor([lr, betas[0], betas[1], eps, wd, 0, 0, 0], dtype=torch.float32, device=dev)
    steps = [torch.zeros((), dtype=torch.int64, device=dev) for _ in range(world)]
    ldc = ops.logits_ld(Cc)
    ref = {}
    rows = torch.arange(Cc, device=dev)
    for step in range(3):
        xs, dls, dbps, gw, gb = [], [], [], None, None
        for r in range(world):
            x = (torch.randn(B, D, generator=g) * 0.5).to(torch.bfloat16).to(dev)
            dl = torch.zeros((B, ldc), dtype=torch.bfloat16)
            dl[:, :Cc] = (torch.randn(B, Cc, generator=g) * 1e-2).to(torch.bfloat16)
            dl = dl.to(dev)
            dbp = torch.randn(3, ldc, generator=g).to(dev) if step == 0 else None
            xs.append(x); dls.append(dl); dbps.append(dbp)
            dW, db = ops.head_backward(dl, x, Cc, D, scale=0.5, db_partials=dbp, schedule="streamk")  # the push mode's split
            gw = dW if gw is None else gw + dW
            gb = db if gb is None else gb + db
        gw, gb = gw * (1.0 / world), gb * (1.0 / world)
        want_w, want_b = _torch_adamw(w0, b0, gw, gb, step, lr, betas, eps, wd, ref)
        for r in range(world):
            ops.head_backward(dls[r], xs[r], Cc, D, scale=0.5, db_partials=dbps[r], push=(ctrl_ptrs[r], ready, stage_ptrs, r))
        for r in range(world):
            mw, vw, mb, vb = moments[r]
            ops.grad_exchange_adamw(w16_ptrs, bias_ptrs, ctrl_ptrs, 0, 0, 0, stage_ptrs[r], r, Cc, D, masters[r][0],
                                    masters[r][1], mw, vw, mb, vb, hyper, steps[r], no_wait=True)
        torch.cuda.synchronize()
        want16 = want_w.to(torch.bfloat16)
        for r in range(world):
            own = (rows // 128) % world == r
            assert int(steps[r].item()) == step + 1
            torch.testing.assert_close(masters[r][0][own], want_w[own], rtol=2e-6, atol=1e-8)
            torch.testing.assert_close(masters[r][1][own], want_b[own], rtol=2e-6, atol=1e-8)
            if world > 1:  # rows of other ranks' blocks were not touched
                assert torch.equal(masters[r][0][~own], w0[~own])
            o = ctrl_words + n_stage
            w16 = bufs[r][o: o + Cc * D // 2].view(torch.bfloat16).view(Cc, D)
            bias = bufs[r][o + n_w16: o + n_w16 + Cpad]
            # the operand is the bf16 rounding of the master: equal to the reference's rounding except where the
            # two fp32 values straddle a rounding boundary (then one bf16 ulp apart)
            diff = (w16.float() - want16.float()).abs()
            assert float((diff > 0).float().mean()) < 2e-3 and float(diff.max()) <= float(want16.float().abs().max()) * 2 ** -7
            torch.testing.assert_close(bias[:Cc], want_b, rtol=2e-6, atol=1e-8)
            assert torch.all(bias[Cc:] == 0)
            for q in range(world):  # every copy identical, bit for bit
                assert torch.equal(bufs[q][o: o + n_w16 + n_bias], bufs[0][o: o + n_w16 + n_bias])


def test_sharded_adamw_module_matches_torch_adamw():
    """SuperGuessr + model.sharded_adamw() on one GPU against the same model trained with torch.optim.AdamW: three
    steps of the smoothed-label loss; losses, master weights and the bf16 operand must agree."""
    import copy

    from geoguessr_ai_b200 import SuperGuessr, synth

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    a = SuperGuessr(None, panorama=True, should_smooth_labels=True, embed_dim=256).to(dev).train()
    b = copy.deepcopy(a)
    C = a.num_cells
    opt_a = a.sharded_adamw(lr=2e-3, betas=(0.9, 0.98), weight_decay=0.02)
    opt_b = torch.optim.AdamW(b.parameters(), lr=2e-3, betas=(0.9, 0.98), weight_decay=0.02)
    sched = torch.optim.lr_scheduler.StepLR(opt_a, step_size=1, gamma=0.5)  # schedulers drive param_groups as usual
    sched_b = torch.optim.lr_scheduler.StepLR(opt_b, step_size=1, gamma=0.5)
    for step in range(3):
        emb, _, _, labels = synth.head_inputs(96, 256, C, seed=10 + step)
        emb, labels = emb.to(dev), labels.to(dev)
        clf = torch.zeros(96, dtype=torch.long, device=dev)
        out_a = a(embedding=emb, labels=labels, labels_clf=clf)
        out_b = b(embedding=emb, labels=labels, labels_clf=clf)
        assert torch.allclose(out_a.loss, out_b.loss, rtol=1e-4), (step, out_a.loss.item(), out_b.loss.item())
        opt_a.zero_grad(); opt_b.zero_grad()
        out_a.loss.backward(); out_b.loss.backward()
        assert a.cell_layer.weight.grad is None  # no gradient is materialised
        opt_a.step(); opt_b.step()
        sched.step(); sched_b.step()
        torch.testing.assert_close(a.cell_layer.weight.data, b.cell_layer.weight.data, rtol=5e-6, atol=1e-5)
        torch.testing.assert_close(a.cell_layer.bias.data, b.cell_layer.bias.data, rtol=5e-6, atol=1e-5)
    with pytest.raises(RuntimeError):  # one backward per step
        a(embedding=emb, labels=labels, labels_clf=clf).loss.backward()
        a(embedding=emb, labels=labels, labels_clf=clf).loss.backward()
    opt_a.step()
    a.eval(); b.eval()
    ya, yb = a(embedding=emb, labels=labels, labels_clf=clf), b(embedding=emb, labels=labels, labels_clf=clf)
    assert torch.equal(ya.top5_geocells.indices[:, 0], yb.top5_geocells.indices[:, 0]) or \
        float((ya.top5_geocells.indices[:, 0] == yb.top5_geocells.indices[:, 0]).float().mean()) > 0.98


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
