#!/usr/bin/env python
"""Benchmark of the post-encoder geolocation hot path (BASELINE.json contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload train|retrieve]

Default workload = BASELINE.json configs[1]: geocell-head training step with the haversine
label-smoothed CE, batch 4096 per GPU, D = 1024, C = 12 647, bf16 operands, 1 x B200.  A "step" is
what the reference's trainer does per batch (main_coordinator_idun_s3.py:393-424): zero_grad ->
SuperGuessr.forward (fusion, head, top-5, smoothed CE) -> loss.backward() (dW, db) -> [N > 1:
NCCL all-reduce of the gradients] -> AdamW step.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

C_CELLS = 12647
TRAIN = dict(B=4096, D=1024, V=4, k=5)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]),
                    tf_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def profile_traffic(launcher):
    """dram read + write bytes per launch of the launcher's main kernel, from the committed ncu capture
    (profiles/traffic.json, written by tools/ncu_traffic.py from `ncu --set full`); None if not captured."""
    path = os.path.join(REPO, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(launcher, {}).get("dram_bytes")
    except (OSError, ValueError):
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        top = sorted(sm)[len(sm) // 2:] if sm else []  # samples under load = upper half
        return {"sm_mhz": statistics.median(top) if top else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def make_batches(n, B, D, V, seed0):
    from geoguessr_ai_b200 import synth

    out = []
    for i in range(n):
        emb, _, _, labels = synth.head_inputs(B, D, 8, V=V, seed=seed0 + i)
        out.append((emb, labels))
    return out


def cpu_reference_train(steps, warmup, B, D, V, threads=None):
    """The reference's own CPU path for the same step, restated by the oracle (eager PyTorch on the
    host cores): forward + autograd backward + AdamW.  Returns (samples/s, ms/step, cores)."""
    from geoguessr_ai_b200 import synth
    from geoguessr_ai_b200.geocells import load_packaged_centroids
    from oracle import super_guessr_oracle as sgo

    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    cent = load_packaged_centroids()
    emb, W, b, labels = synth.head_inputs(B, D, C_CELLS, V=V, seed=1)
    w = W.clone().requires_grad_(True)
    bb = b.clone().requires_grad_(True)
    opt = torch.optim.AdamW([w, bb], lr=1e-4)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        out = sgo.forward(emb, w, bb, cent, labels, None)  # the smoothed loss never reads labels_clf
        out.loss.backward()
        opt.step()
        float(out.loss.detach())
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    return B / (ms / 1e3), ms, torch.get_num_threads()


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    cfg = TRAIN
    steps, warmup = min(args.steps, 5), min(args.warmup, 1)
    val, ms, cores = cpu_reference_train(steps, warmup, cfg["B"], cfg["D"], cfg["V"])
    line = {
        "impl": "reference", "metric": "head-train samples/s", "value": val, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": train_config(args.gpus, cfg),
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} full steps of batch {cfg['B']} (oracle = eager CPU PyTorch restatement of "
                                   "models/super_guessr.py forward + autograd backward + AdamW)"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def train_config(n_gpus, cfg):
    return {"workload": "BASELINE configs[1]: geocell head training step, haversine label-smoothed CE + backward + "
                        "AdamW, synthetic CLIP ViT-L/14 embeddings",
            "batch_per_gpu": cfg["B"], "global_batch": cfg["B"] * n_gpus, "embed_dim": cfg["D"], "headings": cfg["V"],
            "geocells": C_CELLS, "num_candidates": cfg["k"], "parallelism": f"dp{n_gpus}", "cuda_graph": None,
            "grad_allreduce": None,
            "l2": "working set per step (~0.6 GB: embeddings, W, logits, dlogits, dW, AdamW state) exceeds the 126 MB "
                  "L2; input batches rotate over 3 resident buffers"}


def run_b200_train(args):
    import torch.distributed as dist

    import geoguessr_ai_b200 as gg
    from geoguessr_ai_b200 import ops
    from geoguessr_ai_b200.geocells import load_packaged_centroids

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (sm_100a); there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cfg = TRAIN
    B, D, V = cfg["B"], cfg["D"], cfg["V"]
    K, Wm = args.steps, max(args.warmup, 3)

    cent = load_packaged_centroids()
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        model = gg.SuperGuessr(None, panorama=True, should_smooth_labels=True, embed_dim=D, centroids=cent,
                               num_candidates=cfg["k"]).to(dev)
    torch.manual_seed(0)
    with torch.no_grad():
        model.cell_layer.weight.uniform_(-1 / D ** 0.5, 1 / D ** 0.5)
        model.cell_layer.bias.uniform_(-1 / D ** 0.5, 1 / D ** 0.5)
    model.train()
    if world > 1:  # gradients are averaged inside backward(), range by range, overlapped with the dW GEMM
        model.enable_data_parallel(chunks=args.dp_chunks,
                                   comm_dtype=torch.bfloat16 if args.dp_bf16 else None,
                                   comm="nccl" if args.dp_bf16 else args.dp_comm)
    params = [model.cell_layer.weight, model.cell_layer.bias]
    use_graph = not args.no_graph
    opt = torch.optim.AdamW(params, lr=1e-4, fused=True, capturable=use_graph)

    host = make_batches(3, B, D, V, seed0=100 + 10 * rank)
    host = [(e.pin_memory(), l.pin_memory()) for e, l in host]
    resident = [(e.to(dev), l.to(dev)) for e, l in host]
    dummy_clf = torch.zeros(B, dtype=torch.int64, device=dev)  # unused by the smoothed loss (as in the reference)

    def step(emb, labels):
        opt.zero_grad(set_to_none=True)
        out = model(embedding=emb, labels=labels, labels_clf=dummy_clf)
        out.loss.backward()
        opt.step()
        return out.loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # One CUDA graph per input buffer: the ~25 launches of a step (ours, the fused AdamW, the NCCL
    # all-reduce) replay without host work in between.  Falls back to eager launches if capture fails.
    graphs = {}

    def capture(key, emb, labels):
        g = torch.cuda.CUDAGraph()
        pool = next(iter(graphs.values()))[0].pool() if graphs else None
        with torch.cuda.graph(g, pool=pool):
            loss = step(emb, labels)
        graphs[key] = (g, loss)

    def run(key, emb, labels):
        if key in graphs:
            g, loss = graphs[key]
            g.replay()
            return loss
        return step(emb, labels)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):  # eager warm-up off the default stream (required before capture)
        for i in range(3):
            step(*resident[i % 3])
    torch.cuda.current_stream().wait_stream(side)
    barrier()
    if use_graph:
        try:
            for i in range(3):
                capture(("res", i), *resident[i])
        except Exception as e:  # noqa: BLE001
            graphs.clear()
            use_graph = False
            torch.cuda.synchronize()
            if rank == 0:
                print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); eager launches", file=sys.stderr)

    # ---------------- device-resident throughput (`value`)
    for i in range(Wm):
        run(("res", i % 3), *resident[i % 3])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        loss = run(("res", i % 3), *resident[i % 3])
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / K
    value = world * B / (ms_step / 1e3)
    final_loss = loss.item()

    # ---------------- per-launcher CUDA-event timing over a second identical timed region
    # Eager launches; a ~1 ms device-side spin queued ahead of every step lets the host run a whole step
    # ahead, so the events bracket device time only (no launch gaps inside the brackets).
    ops.enable_timing(True)
    barrier()
    n_timed = min(K, 20)
    for i in range(n_timed):
        torch.cuda._sleep(2_000_000)
        step(*resident[i % 3])
    tms = ops.timing_ms()
    ops.enable_timing(False)
    per = {k: sum(v) / n_timed for k, v in tms.items()}  # per step (a launcher may run more than once: --dp-chunks)

    # ---------------- end to end from host buffers (`e2e`): H2D of the batch + D2H of the loss every step
    copy_stream = torch.cuda.Stream()
    bufs = [(torch.empty_like(resident[0][0]), torch.empty_like(resident[0][1])) for _ in range(2)]
    evs = [torch.cuda.Event(), torch.cuda.Event()]

    def issue_copy(i):
        with torch.cuda.stream(copy_stream):
            bufs[i % 2][0].copy_(host[i % 3][0], non_blocking=True)
            bufs[i % 2][1].copy_(host[i % 3][1], non_blocking=True)
            evs[i % 2].record(copy_stream)

    if use_graph:
        try:
            for i in range(2):
                capture(("e2e", i), *bufs[i])
        except Exception:  # noqa: BLE001
            for i in range(2):
                graphs.pop(("e2e", i), None)
            torch.cuda.synchronize()

    def e2e_loop(n):
        issue_copy(0)
        for i in range(n):
            torch.cuda.current_stream().wait_event(evs[i % 2])
            l = run(("e2e", i % 2), *bufs[i % 2])
            if i + 1 < n:
                issue_copy(i + 1)  # next batch crosses PCIe while this step computes
            l.item()  # device -> host read of the step's result

    Ke = max(3, min(K, 20))
    e2e_loop(3)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(Ke)
    torch.cuda.synchronize()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3) / Ke
    e2e_val = world * B / (ms_e2e / 1e3)
    h2d = host[0][0].numel() * 4 + host[0][1].numel() * 4

    if rank != 0:
        finish(world)
        return

    # ---------------- roofline of the dominant launcher
    peaks = measured_peaks()
    C = C_CELLS
    flops_gemm = 2.0 * B * C * D
    algo = {
        "gg_head_fwd": ("tensor", flops_gemm, "TFLOP/s"),
        "gg_head_bwd": ("tensor", flops_gemm, "TFLOP/s"),
        "gg_hav_ce_fwd_bwd": ("hbm", B * C * (2 + 2) + 8 * B + 12 * C, "GB/s"),
        "gg_fuse_headings": ("hbm", B * D * (4 * V + 2), "GB/s"),
        "gg_prepare_head_weights": ("hbm", C * D * (4 + 2) + 8 * C, "GB/s"),
        "gg_fuse_and_prepare": ("hbm", B * D * (4 * V + 2) + C * D * (4 + 2) + 8 * C, "GB/s"),
    }
    kernels = {}
    for name, ms in per.items():
        if name in algo:
            bound, work, unit = algo[name]
            ach = work / (ms * 1e-3) / (1e12 if bound == "tensor" else 1e9)
            peak = peaks["tf_sustained"] if bound == "tensor" else peaks["hbm"]
            kernels[name] = {"ms": ms, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak}
        else:
            kernels[name] = {"ms": ms}
    dom = max((k for k in kernels if "frac" in kernels[k]), key=lambda k: kernels[k]["ms"])
    roof = dict(kernels[dom])
    roof.pop("ms")
    roof.update({"kernel": dom, "ms_per_launch": kernels[dom]["ms"], "traffic": profile_traffic(dom),
                 "peak_source": f"MEASURED_PEAKS.json ({peaks['source']}; "
                                + ("bf16_tflops_sustained: kernel timed inside the step" if roof["bound"] == "tensor"
                                   else "hbm_gbs") + ")"})

    cpu = None
    if not args.no_cpu:
        cval, cms, cores = cpu_reference_train(2, 1, B, D, V)
        cpu = {"value": cval, "unit": "samples/s", "cores": cores, "kind": "port", "ms_per_step": cms,
               "sample": f"2 full steps of batch {B} after 1 warm-up (oracle: eager CPU PyTorch restatement of the "
                         "reference forward + autograd backward + AdamW)"}

    grad_comm = None
    if world > 1:
        if model._dp["symm"] is not None:
            grad_comm = (("NVSwitch multicast (multimem.ld_reduce / multimem.st)" if model._dp["symm"]["kind"] == "nvls"
                          else "peer load/store two-shot (rank-order sums)")
                         + " all-reduce (avg, fp32) of [dW | db] in symmetric memory over NVLink, one kernel per rank "
                           f"and geocell range between two symmetric-memory barriers; {args.dp_chunks} range(s), each "
                           "exchanged while the next one's dW GEMM runs")
        else:
            grad_comm = (f"nccl avg, {'bf16' if args.dp_bf16 else 'fp32'}, {args.dp_chunks} geocell ranges overlapped "
                         "with the dW GEMM")
    # fusion + weight cast, head_fwd + merge, label vectors + row statistics, loss stream kernel, dW GEMM [, gradient exchange]
    launches_per_step = 1 + 2 + 2 + 1 + 1 + (args.dp_chunks - 1 + args.dp_chunks if world > 1 and model._dp["symm"] is not None else 0)
    line = {
        "metric": "head-train samples/s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K,
        "warmup": Wm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": dict(train_config(world, cfg), cuda_graph=use_graph,
                                            grad_allreduce=grad_comm), "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "samples/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "steps": Ke},
        "gpu_launches": launches_per_step * K, "roofline": roof, "kernels": kernels,
        "kernels_timing": "CUDA events around every launcher in a second, eager pass over the same steps with nothing "
                          "running beside the timed launcher (the label statistics, which the timed region runs "
                          "underneath the head GEMM on a side stream, run serially there); a launcher may hold a small "
                          "helper kernel (gg_head_fwd: GEMM + merge)",
        "cpu_baseline": cpu, "loss": final_loss,
    }
    print(json.dumps(line), flush=True)
    finish(world)


# ------------------------------------------------------------------------------------------------
# Serving workload (BASELINE configs[2] / [4]): SuperGuessr serving forward -> ProtoRefiner on a geocell-sharded
# prototype bank.  Not the default bench line (configs[1] is); run with --workload infer.
INFER = dict(B=65536, D=1024, V=4, k=5)


def infer_config(n_gpus, cfg, P, B):
    return {"workload": f"BASELINE configs[{2 if P <= 1_000_000 else 4}]: SuperGuessr serving + ProtoRefiner retrieval vs "
                        f"{P} synthetic prototypes sharded by geocell, top-{cfg['k']} cells",
            "queries_per_batch": B, "prototypes": P, "embed_dim": cfg["D"], "headings": cfg["V"], "geocells": C_CELLS,
            "parallelism": f"queries split {n_gpus}-way in front of a bank sharded {n_gpus}-way by geocell",
            "l2": "every batch streams the whole local bank shard (>= 0.25 GB) and 1 GB of embeddings: far beyond the "
                  "126 MB L2"}


def make_local_bank(P, D, rank, world, dev, cent):
    """Synthetic CSR bank, rows of this rank's geocell range only (generated on the device)."""
    from geoguessr_ai_b200 import shard_cells, synth

    sizes = synth.cell_sizes(C_CELLS, P, seed=0, mode="skewed")
    off = np.zeros(C_CELLS + 1, dtype=np.int64)
    np.cumsum(sizes, out=off[1:])
    lo, hi = shard_cells(off, world)[rank]
    p0, p1 = int(off[lo]), int(off[hi])
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    bank = torch.empty((p1 - p0, D), dtype=torch.bfloat16, device=dev)
    for a in range(0, p1 - p0, 1 << 20):
        b = min(p1 - p0, a + (1 << 20))
        bank[a:b] = torch.randn((b - a, D), device=dev, generator=gen).to(torch.bfloat16)
    cell_of = torch.repeat_interleave(torch.arange(lo, hi, device=dev), torch.from_numpy(sizes[lo:hi]).to(dev))
    coords = cent.to(dev)[cell_of] + (torch.rand((p1 - p0, 2), device=dev, generator=gen) - 0.5)
    return torch.from_numpy(off.astype(np.int32)), bank, coords.float(), (lo, hi, p0, p1)


def cpu_reference_infer(model_w, model_b, cent, emb, refiner, nq):
    """The reference's serving path restated by the oracle on `nq` queries: eager CPU head + the Python
    (query x candidate) loop of ProtoRefiner.forward.  Prototypes of the cells those queries touch are copied
    back from the device bank.  Returns (queries/s, cores)."""
    from oracle import proto_refiner_oracle as pro
    from oracle import super_guessr_oracle as sgo

    torch.set_num_threads(os.cpu_count() or 1)
    e = emb[:nq].float().cpu()
    W, b = model_w.detach().float().cpu(), model_b.detach().float().cpu()
    out = sgo.forward(e, W, b, cent, None, torch.zeros(nq, dtype=torch.int64))  # warm-up + candidates
    cells = sorted(set(out.top5_geocells.indices.flatten().tolist()))
    off = refiner.cell_off.cpu().tolist()
    protos, coords = [None] * C_CELLS, [None] * C_CELLS
    for c in cells:
        if refiner.cell_lo <= c < refiner.cell_hi:
            a, z = off[c - refiner.cell_lo], off[c - refiner.cell_lo + 1]
            if z > a:
                protos[c] = refiner.bank[a:z].float().cpu()
                coords[c] = refiner.bank_coords[a:z].cpu()
    t0 = time.perf_counter()
    out = sgo.forward(e, W, b, cent, None, torch.zeros(nq, dtype=torch.int64))
    pro.forward(e, out.preds_LLH, out.top5_geocells.indices, out.top5_geocells.values.detach(), protos, coords, topk=5)
    dt = time.perf_counter() - t0
    return nq / dt, torch.get_num_threads()


def run_b200_infer(args):
    import contextlib
    import io

    import torch.distributed as dist

    import geoguessr_ai_b200 as gg
    from geoguessr_ai_b200 import ops
    from geoguessr_ai_b200.geocells import load_packaged_centroids

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (sm_100a); there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cfg = INFER
    D, V, k = cfg["D"], cfg["V"], cfg["k"]
    B, P = args.batch, args.protos
    assert B % world == 0
    Bl = B // world
    K, Wm = args.steps, max(args.warmup, 3)

    cent = load_packaged_centroids()
    with contextlib.redirect_stdout(io.StringIO()):
        model = gg.SuperGuessr(None, panorama=True, serving=True, embed_dim=D, centroids=cent, num_candidates=k).to(dev)
    torch.manual_seed(0)  # same head on every rank
    with torch.no_grad():
        model.cell_layer.weight.uniform_(-1 / D ** 0.5, 1 / D ** 0.5)
        model.cell_layer.bias.uniform_(-1 / D ** 0.5, 1 / D ** 0.5)
    model.eval()
    off, bank, coords, (lo, hi, p0, p1) = make_local_bank(P, D, rank, world, dev, cent)
    refiner = gg.ProtoRefiner(topk=k, protos="bank", bank=(off, bank, coords), shard=(rank, world),
                              split_queries=world > 1, bank_is_local=True, report_changed=False, device=dev)
    del bank, coords

    gen = torch.Generator(device=dev)
    gen.manual_seed(77 + rank)
    resident = [torch.randn((Bl, V, D), device=dev, generator=gen) for _ in range(2)]
    host = [r.cpu().pin_memory() for r in resident[:1]]

    def step(emb):
        llh, topk, _ = model(embedding=emb)
        _, r_llh, r_cell = refiner(emb, llh, topk.indices, topk.values)
        return r_llh, r_cell

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    for i in range(Wm):
        step(resident[i % 2])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        out = step(resident[i % 2])
    e1.record()
    barrier()
    ms_step = max_over_ranks(e0.elapsed_time(e1)) / K
    clocks = sampler.stop() if rank == 0 else None
    value = B / (ms_step / 1e3)

    ops.enable_timing(True)
    barrier()
    for i in range(min(K, 5)):
        torch.cuda._sleep(40_000_000)
        step(resident[i % 2])
    tms = ops.timing_ms()
    ops.enable_timing(False)
    per = {kk: sum(v) / len(v) for kk, v in tms.items()}
    calls = {kk: len(v) // min(K, 5) for kk, v in tms.items()}

    # end to end: the batch crosses PCIe from pinned memory, results come back to the host
    copy_stream = torch.cuda.Stream()
    bufs = [torch.empty_like(resident[0]) for _ in range(2)]
    evs = [torch.cuda.Event(), torch.cuda.Event()]
    res_host = (torch.empty((Bl, 2), dtype=torch.float32).pin_memory(), torch.empty((Bl,), dtype=torch.int64).pin_memory())

    def issue_copy(i):
        with torch.cuda.stream(copy_stream):
            bufs[i % 2].copy_(host[0], non_blocking=True)
            evs[i % 2].record(copy_stream)

    def e2e_loop(n):
        issue_copy(0)
        for i in range(n):
            torch.cuda.current_stream().wait_event(evs[i % 2])
            r_llh, r_cell = step(bufs[i % 2])
            if i + 1 < n:
                issue_copy(i + 1)
            res_host[0].copy_(r_llh, non_blocking=True)
            res_host[1].copy_(r_cell, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    Ke = max(3, min(K, 10))
    e2e_loop(2)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(Ke)
    torch.cuda.synchronize()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3) / Ke
    h2d = host[0].numel() * 4
    d2h = Bl * (8 + 8)

    if rank != 0:
        finish(world)
        return

    peaks = measured_peaks()
    P_local = p1 - p0
    algo = {
        "gg_head_fwd": ("tensor", 2.0 * Bl * C_CELLS * D, "TFLOP/s"),
        # bank shard read once + grouped queries written and read + records
        "gg_proto_retrieve": ("hbm", P_local * D * 2.0 + 3.0 * B * k * D * 2 + B * k * 16, "GB/s"),
        "gg_fuse_headings": ("hbm", Bl * D * (4 * V + 2.0), "GB/s"),
        "gg_proto_refine": ("hbm", (world * k * 16 + k * 12 + 24.0) * Bl, "GB/s"),
    }
    kernels = {}
    for name, ms in per.items():
        n = max(1, calls.get(name, 1))
        if name in algo:
            bound, work, unit = algo[name]
            ach = work / (ms * 1e-3) / (1e12 if bound == "tensor" else 1e9)
            peak = peaks["tf_sustained"] if bound == "tensor" else peaks["hbm"]
            kernels[name] = {"ms": ms, "calls_per_step": n, "bound": bound, "achieved": ach, "peak": peak, "unit": unit,
                             "frac": ach / peak}
        else:
            kernels[name] = {"ms": ms, "calls_per_step": n}
    dom = max((kk for kk in kernels if "frac" in kernels[kk]), key=lambda kk: kernels[kk]["ms"] * kernels[kk]["calls_per_step"])
    roof = {kk: v for kk, v in kernels[dom].items() if kk not in ("ms", "calls_per_step")}
    roof.update({"kernel": dom, "ms_per_launch": kernels[dom]["ms"], "traffic": profile_traffic(dom),
                 "peak_source": f"MEASURED_PEAKS.json ({peaks['source']})"})
    cpu = None
    if not args.no_cpu:
        nq = 256
        cval, cores = cpu_reference_infer(model.cell_layer.weight, model.cell_layer.bias, cent, resident[0], refiner, nq)
        cpu = {"value": cval, "unit": "queries/s", "cores": cores, "kind": "port",
               "sample": f"{nq} queries of the same batch through the oracle (eager CPU head + the reference's Python "
                         "(query x candidate) refiner loop); prototypes of the touched cells copied from the device bank"
                         + ("" if world == 1 else "; cells owned by other ranks count as missing")}
    launches = 2 + 2 + 1 + 6 + 1  # fuse, head (GEMM + merge), refiner fuse, retrieval (memset + 5 kernels), refine
    line = {
        "metric": "geolocation queries/s", "value": value, "unit": "queries/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": infer_config(world, cfg, P, B), "clocks": clocks,
        "e2e": {"value": B / (ms_e2e / 1e3), "unit": "queries/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": Ke},
        "gpu_launches": launches * K, "roofline": roof, "kernels": kernels, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    finish(world)


def finish(world):
    """Multi-rank exit: tearing NCCL communicators down while CUDA graphs that captured collectives are alive can
    block; everything is measured and printed by now, so leave without running destructors."""
    if world > 1:
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    # NCCL prints its version banner on stdout when NCCL_DEBUG=VERSION; the bench contract is ONE JSON line
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "infer"],
                    help="train = BASELINE configs[1] (the bench line); infer = serving + sharded prototype retrieval")
    ap.add_argument("--protos", type=int, default=1_000_000, help="infer: total prototypes (configs[2]: 1e6, configs[4]: 1e7)")
    ap.add_argument("--batch", type=int, default=65536, help="infer: queries per batch over all GPUs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiler runs only)")
    ap.add_argument("--dp-chunks", type=int, default=1, help="geocell ranges of the overlapped dW GEMM + all-reduce (N > 1)")
    ap.add_argument("--dp-comm", default="auto", choices=["auto", "nvls", "p2p", "nccl"],
                    help="gradient exchange for N > 1: own kernel over symmetric memory (NVSwitch multicast or peer "
                         "loads/stores) or NCCL all-reduce")
    ap.add_argument("--dp-bf16", action="store_true", help="all-reduce the gradients in bf16 (opt-in, N > 1)")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying CUDA graphs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "infer":
        run_b200_infer(args)
    else:
        run_b200_train(args)


if __name__ == "__main__":
    main()
