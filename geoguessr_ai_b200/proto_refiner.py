"""ProtoRefiner: drop-in for the reference's proto-net refinement (models/proto_refiner.py:30-237)
running on the sm_100a retrieval + refinement kernels.

Kept from the reference: constructor keywords, ``forward(embedding, initial_preds, candidate_cells,
candidate_probs=None) -> (loss, preds_LLH (B,2), preds_geocell (B,))``, the frozen ``temperature`` /
``geo_scaling`` parameters (state-dict keys, :117-118), the missing-cell sentinel (-100000, (0,0)),
the un-stabilised temperature softmax, the 1000 km guard and the "Changed geocell predictions"
report (:231-233).  As in the executable reference the similarity is NEGATIVE EUCLIDEAN distance
(``_euclidean_distance``; ``_cosine_similarity`` is never called), and a prototype's coordinates
are the ones stored with it (the ``count == 0`` branch of ``_within_cluster_refinement``, the only
one that can run: the other dereferences a ``self.dataset`` that is never assigned).

The reference keeps the bank as a python list of per-cell HF datasets and copies a cell's
prototypes host->device for every (query, candidate).  Here the bank lives in HBM once, sorted by
geocell: ``bank`` (P, D) bf16, CSR ``cell_off`` (C+1), ``coords`` (P, 2), ``sqnorm`` (P).  With
``shard=(rank, world)`` each rank holds a contiguous geocell range (balanced by prototype count),
fills the candidates it owns and the per-pair records are merged with one all-gather.

Out of scope here: building prototypes from images (``protos=None`` in the
reference starts the offline embedding job, proto_refiner.py:89-103,313-345).
"""
from __future__ import annotations

import os
from typing import Sequence

import numpy as np
import torch
from torch import Tensor, nn
from torch.nn.parameter import Parameter

from . import ops

PROTO_PATH = "data/geocells/proto_df.csv"


def shard_cells(cell_off_cpu: np.ndarray, world: int):
    """Contiguous geocell ranges [lo, hi) per rank, balanced by prototype count."""
    C = len(cell_off_cpu) - 1
    P = int(cell_off_cpu[-1])
    bounds = [0]
    for r in range(1, world):
        target = P * r / world
        b = int(np.searchsorted(cell_off_cpu, target, side="left"))
        bounds.append(min(max(b, bounds[-1]), C))
    bounds.append(C)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


class ProtoRefiner(nn.Module):
    def __init__(
        self,
        topk: int = 5,
        max_refinement: int = 1000,
        temperature: float = 1.6,
        proto_path: str = PROTO_PATH,
        protos=None,
        verbose: bool = False,
        clip_db_path: str = "data/sqlite/clip/dataset.sqlite",
        tinyvit_db_path: str = "data/sqlite/tinyvit/dataset.sqlite",
        backend: str = "clip",
        *,
        coords: Sequence | None = None,
        bank: tuple | None = None,
        shard: tuple | None = None,
        process_group=None,
        split_queries: bool = False,
        bank_is_local: bool = False,
        report_changed: bool = True,
        precision: str = "bf16",
        metric: str = "l2",
        images: tuple | None = None,
        device="cuda",
    ):
        """Reference arguments (proto_refiner.py:33-62) plus keyword-only ways to hand over a bank:

        protos + coords: python lists with one ``(P_c, D)`` tensor (or None) and one ``(P_c, 2)``
            (lng, lat) tensor per geocell -- the reference's in-memory representation.
        protos="load": per-cell HF datasets under ``data/geocells/protos/proto_{i}`` (:104-112), rows
            carrying ``embedding``, ``centroid_lng``, ``centroid_lat``.
        bank=(cell_off, bank, coords): the CSR form directly (any float dtype; stored as bf16).
        shard=(rank, world): keep only this rank's geocell range; ``process_group`` is the
            torch.distributed group used for the exchange (default group if None).
        split_queries: with a sharded bank, every rank passes a DIFFERENT, equally sized slice of the query
            batch (data-parallel serving in front of the refiner).  forward() then all-gathers the fused bf16
            queries + candidate lists, retrieves on the local shard for all of them, and returns each rank
            the records of its own queries with one all-to-all.  False: every rank passes the same batch
            (records merged with one all-gather).
        bank_is_local: ``bank`` / ``coords`` already hold only this rank's rows (``cell_off`` stays global).
        precision: "bf16" (prototypes and queries rounded to bf16, fp32 accumulate) or "bf16x3" (hi/lo split
            operands, three products: fp32-faithful scores for fp32 banks such as the reference's CLIP prototypes;
            3x the bank bytes).
        metric: "l2" = the executable reference (``-cdist``, :190); "cosine" = the reference's unused
            ``_cosine_similarity`` (:347-362), opt-in.
        images=(member_off, image_bank, image_coords): opt-in within-cluster refinement: the member
            images of every prototype, sorted by prototype -- ``member_off`` (P+1) over the GLOBAL prototype order of
            ``bank``, ``image_bank`` (L, D), ``image_coords`` (L, 2).  After the nearest prototype of a candidate cell
            is found, the nearest member image of that cluster supplies the coordinates (a cluster without images keeps
            its own: the ``count == 0`` branch, :251-252).  This is what ``_within_cluster_refinement`` (:239-269) is
            meant to do; as written it cannot run (``self.dataset`` is never assigned) and takes the farthest member.
        """
        super().__init__()
        if precision not in ("bf16", "bf16x3"):
            raise ValueError("precision must be 'bf16' or 'bf16x3'")
        if metric not in ops.METRICS:
            raise ValueError("metric must be 'l2' or 'cosine'")
        self.precision, self.metric = precision, metric
        self.topk = topk
        self.max_refinement = max_refinement
        self.verbose = verbose
        self.report_changed = report_changed
        self.process_group = process_group
        self.split_queries = bool(split_queries)
        self.temperature = Parameter(torch.tensor(float(temperature)), requires_grad=False)
        self.geo_scaling = Parameter(torch.tensor(20.0), requires_grad=False)
        self._temperature_host = float(temperature)

        if bank is not None:
            cell_off, mat, xy = bank
        elif isinstance(protos, (list, tuple)):
            if coords is None:
                raise ValueError("protos given as a list needs the matching coords list")
            cell_off, mat, xy = self._csr_from_lists(protos, coords)
        elif isinstance(protos, str):
            cell_off, mat, xy = self._csr_from_disk(proto_path)
        else:
            raise NotImplementedError(
                "protos=None asks the reference to BUILD prototypes from images (embedders + S3, "
                "proto_refiner.py:89-103).  With stored embeddings that job is "
                "geoguessr_ai_b200.proto_builder.build_prototype_bank(embeddings, proto_df, num_cells) -> pass its "
                "result as bank=(cell_off, bank, coords); or pass protos='load' / protos=[...]+coords=[...].")
        self.has_images = False
        self._install_bank(cell_off, mat, xy, shard, device, bank_is_local)
        self._install_images(images, device)

    # ---- bank construction ------------------------------------------------------------------
    @classmethod
    def from_bank(cls, cell_off, bank, coords, **kw):
        return cls(protos="bank", bank=(cell_off, bank, coords), **kw)

    @staticmethod
    def _csr_from_lists(protos, coords):
        sizes = [0 if p is None else int(p.shape[0]) for p in protos]
        off = np.zeros(len(sizes) + 1, dtype=np.int64)
        np.cumsum(sizes, out=off[1:])
        mats = [p.float() for p in protos if p is not None and p.shape[0] > 0]
        xys = [torch.as_tensor(c, dtype=torch.float32).reshape(-1, 2)
               for p, c in zip(protos, coords) if p is not None and p.shape[0] > 0]
        D = mats[0].shape[1] if mats else 8
        mat = torch.cat(mats, 0) if mats else torch.zeros((0, D))
        xy = torch.cat(xys, 0) if xys else torch.zeros((0, 2))
        return torch.from_numpy(off.astype(np.int32)), mat, xy

    @staticmethod
    def _csr_from_disk(proto_path):
        from datasets import Dataset  # HF datasets, as in the reference (:108-110)

        root = "data/geocells/protos"
        if os.path.exists(proto_path):
            import pandas as pd

            num = int(pd.read_csv(proto_path)["geocell_index"].astype(int).max()) + 1
        else:
            ids = [int(n.split("_")[1]) for n in os.listdir(root) if n.startswith("proto_")]
            num = max(ids) + 1
        protos, coords = [], []
        for i in range(num):
            try:
                ds = Dataset.load_from_disk(f"{root}/proto_{i}").with_format("torch")
            except FileNotFoundError:
                protos.append(None)
                coords.append(None)
                continue
            protos.append(ds["embedding"].float())
            coords.append(torch.stack([ds["centroid_lng"].float(), ds["centroid_lat"].float()], 1))
        return ProtoRefiner._csr_from_lists(protos, coords)

    def _install_bank(self, cell_off, mat, xy, shard, device, bank_is_local=False):
        cell_off_cpu = torch.as_tensor(cell_off).to(torch.int64).cpu().numpy()
        self.num_geocells = len(cell_off_cpu) - 1
        self.num_protos_total = int(cell_off_cpu[-1])
        self.embed_dim = int(mat.shape[1])
        rank, world = shard if shard is not None else (0, 1)
        self.rank, self.world = int(rank), int(world)
        lo, hi = shard_cells(cell_off_cpu, self.world)[self.rank]
        self.cell_lo, self.cell_hi = lo, hi
        p0, p1 = int(cell_off_cpu[lo]), int(cell_off_cpu[hi])
        self.proto_base = p0
        local_off = torch.from_numpy((cell_off_cpu[lo:hi + 1] - p0).astype(np.int32))
        dev = torch.device(device)
        if bank_is_local:
            assert mat.shape[0] == p1 - p0 and len(xy) == p1 - p0, "local bank does not match this rank's geocell range"
        else:
            mat, xy = mat[p0:p1], xy[p0:p1]
        self._split = self.precision == "bf16x3"
        mat = torch.as_tensor(mat)
        if dev.type == "cuda" and mat.dtype == torch.float32:
            bank16 = ops.cast_bank_bf16(mat.to(dev), split=self._split)  # gg_cast_bf16: no torch arithmetic
        elif self._split:
            hi = mat.to(torch.bfloat16)
            lo = (mat.float() - hi.float()).to(torch.bfloat16)
            bank16 = torch.cat([hi, lo, hi], 1).to(dev).contiguous()
        else:
            bank16 = mat.to(device=dev, dtype=torch.bfloat16).contiguous()
        self.register_buffer("bank", bank16, persistent=False)
        self.register_buffer("bank_coords", torch.as_tensor(xy, dtype=torch.float32).to(dev).contiguous(),
                             persistent=False)
        self.register_buffer("cell_off", local_off.to(dev), persistent=False)
        groups = ops.proto_group_cells(local_off.numpy())
        self.num_groups = len(groups) - 1
        # Query rows: gathered by the TMA engine four at a time (tile::gather4, no cell-ordered copy in HBM) when a
        # work item meets its rows once -- groups of <= 256 prototypes; measured at ~45 ns per gather4 on B200, it
        # loses to one pre-gathered copy when big cells re-read the rows for every 256-prototype unit.
        self.gather4 = self._gather4_default and (p1 - p0) <= 256 * max(self.num_groups, 1)
        self.register_buffer("group_off", torch.from_numpy(groups).to(dev), persistent=False)
        # ||p||^2 needs the CUDA library: computed now on a CUDA device, else on the first forward after .to("cuda")
        self.register_buffer("bank_sqnorm", torch.empty(0, dtype=torch.float32, device=dev), persistent=False)
        self._sqnorm_for = None
        self._last_meta = None
        if dev.type == "cuda":
            self._ensure_sqnorm()

    def _install_images(self, images, device):
        """Second-stage bank: member images sorted by prototype; this rank keeps those of its own prototypes."""
        if images is None:
            return
        member_off, img, img_xy = images
        moff = torch.as_tensor(member_off).to(torch.int64).cpu().numpy()
        assert len(moff) == self.num_protos_total + 1, "member_off must have one entry per prototype of the whole bank + 1"
        p0 = self.proto_base
        p1 = p0 + int(self.bank.shape[0])
        i0, i1 = int(moff[p0]), int(moff[p1])
        self.img_base = i0
        dev = torch.device(device)
        local = torch.from_numpy((moff[p0:p1 + 1] - i0).astype(np.int32))
        img = torch.as_tensor(img)[i0:i1]
        if dev.type == "cuda" and img.dtype == torch.float32:
            img16 = ops.cast_bank_bf16(img.to(dev), split=self._split)
        elif self._split:
            hi = img.to(torch.bfloat16)
            img16 = torch.cat([hi, (img.float() - hi.float()).to(torch.bfloat16), hi], 1).to(dev).contiguous()
        else:
            img16 = img.to(device=dev, dtype=torch.bfloat16).contiguous()
        self.has_images = True
        self.register_buffer("img_bank", img16, persistent=False)
        self.register_buffer("img_coords", torch.as_tensor(img_xy, dtype=torch.float32)[i0:i1].to(dev).contiguous(),
                             persistent=False)
        self.register_buffer("img_off", local.to(dev), persistent=False)
        self.register_buffer("img_group_off", torch.from_numpy(ops.proto_group_cells(local.numpy())).to(dev), persistent=False)
        self.register_buffer("img_sqnorm", torch.empty(0, dtype=torch.float32, device=dev), persistent=False)
        self._img_sqnorm_for = None
        if dev.type == "cuda":
            self._ensure_sqnorm()

    def _ensure_sqnorm(self):
        """``bank_sqnorm`` of the bank where it lives NOW (the reference idiom is ``ProtoRefiner(...).to(device)``,
        inference.py:177: a bank installed on the CPU has no norms until it reaches the GPU)."""
        key = (self.bank.data_ptr(), self.bank.device)
        if self._sqnorm_for != key:
            self.bank_sqnorm = (ops.row_sqnorm_bf16(self.bank, split=self._split) if self.bank.shape[0]
                                else torch.empty(0, dtype=torch.float32, device=self.bank.device))
            self._sqnorm_for = key
        if self.has_images:
            key = (self.img_bank.data_ptr(), self.img_bank.device)
            if self._img_sqnorm_for != key:
                self.img_sqnorm = (ops.row_sqnorm_bf16(self.img_bank, split=self._split) if self.img_bank.shape[0]
                                   else torch.empty(0, dtype=torch.float32, device=self.img_bank.device))
                self._img_sqnorm_for = key

    def __str__(self):
        rep = "ProtoRefiner(\n"
        rep += f"\ttopk\t\t= {self.topk}\n"
        rep += f"\tmax_refinement\t= {self.max_refinement}\n"
        rep += f"\ttemperature\t= {self.temperature.data.item()}\n"
        rep += f"\tgeo_scaling\t= {self.geo_scaling.data.item()}\n"
        rep += ")"
        return rep

    # ---- forward (proto_refiner.py:129-237) -------------------------------------------------
    def _fuse(self, embedding: Tensor):
        # :150-151 mean over headings (shared with the SuperGuessr serving forward that produced the candidates)
        q16, qn = ops.fuse_headings_shared(embedding, split=self._split)
        if q16.shape[1] != self.embed_dim * (3 if self._split else 1):
            raise ValueError(f"embedding dim {q16.shape[1] // (3 if self._split else 1)} != prototype dim {self.embed_dim}")
        return q16, qn

    def _retrieve(self, q16, qn, candidate_cells):
        self._ensure_sqnorm()
        bank = self.bank if self.bank.shape[0] > 0 else None
        rec, meta = ops.proto_retrieve(q16, qn, candidate_cells, self.topk, bank, self.bank_sqnorm, self.bank_coords,
                                       self.cell_off, self.cell_lo, self.cell_hi, self.proto_base,
                                       group_off=self.group_off, metric=self.metric, gather4=self.gather4,
                                       want_meta=True)
        self._last_meta = (meta, q16.shape[0], q16.shape[1], candidate_cells)
        n_local = int(self.bank.shape[0])
        if self.has_images and n_local > 0:
            # stage 1b: the best prototype of every pair is the "cell" of a second retrieval over its member images
            cand2 = ops.proto_record_ids(rec).view(q16.shape[0], self.topk)
            img = self.img_bank if self.img_bank.shape[0] > 0 else None
            rec_img = ops.proto_retrieve(q16, qn, cand2, self.topk, img, self.img_sqnorm, self.img_coords, self.img_off,
                                         self.proto_base, self.proto_base + n_local, self.img_base,
                                         group_off=self.img_group_off, metric=self.metric, gather4=self.gather4)
            ops.proto_take_image_coords(rec, rec_img)
        return rec

    _gather4_default = os.environ.get("GG_RETRIEVE_GATHER4", "1") != "0"

    def last_retrieve_stats(self):
        """Work of the last retrieval on this rank (reads 16 bytes back: synchronises): work items, accumulation
        units and the operand bytes / FLOP they stand for (the launcher's executed traffic, next to the
        algorithmic P*D*2 + pairs*D*2)."""
        if self._last_meta is None:
            return None
        meta, nq, K, cand = self._last_meta
        items, pairs, units, boxes = (int(v) for v in meta.cpu().tolist())
        sizes = (self.cell_off[1:] - self.cell_off[:-1]).to(torch.int64)
        c = cand[:, :self.topk].to(torch.int64) - self.cell_lo
        mine = (c >= 0) & (c < sizes.numel())
        algo_flop = 2.0 * K * float(sizes[c.clamp(0, max(sizes.numel() - 1, 0))][mine].sum().item()) if sizes.numel() else 0.0
        row = K * 2
        # bank boxes of 32 rows up to each group's end + the pairs' rows (4-row gathers / 32-row boxes) per unit + records
        a_rows = (pairs + (2 if self.gather4 else 16) * items) * (units / max(items, 1))
        executed = boxes * 32 * row + a_rows * row + pairs * 16 + nq * self.topk * 16
        if not self.gather4:
            executed += 2 * pairs * row  # the cell-ordered copy: read + write
        return dict(work_items=items, pairs=pairs, units=units, executed_bytes=float(executed),
                    executed_flop=float(units) * 2 * 128 * 256 * K,
                    algorithmic_flop=algo_flop, query_rows=nq)

    def retrieve(self, embedding: Tensor, candidate_cells: Tensor) -> Tensor:
        """Stage 0+1 on this rank's shard: (B*topk, 4) records."""
        q16, qn = self._fuse(embedding)
        return self._retrieve(q16, qn, candidate_cells)

    def forward(self, embedding: Tensor = None, initial_preds: Tensor = None, candidate_cells: Tensor = None,
                candidate_probs: Tensor = None, return_debug: bool = False):
        assert self.topk <= candidate_cells.size(1), (
            '"topk" parameter must be smaller or equal to the number of geocell candidates '
            "passed into the forward function.")
        if not self.bank_coords.is_cuda:
            raise ops._lib.GeoguessrB200Error("ProtoRefiner bank is on the CPU; there is no CPU fallback (use device='cuda')")
        dev = self.bank_coords.device
        embedding, initial_preds, candidate_cells = (t.to(dev) for t in (embedding, initial_preds, candidate_cells))
        if candidate_probs is not None:
            candidate_probs = candidate_probs.to(dev)
        temperature = self._temperature_host  # host copy of the frozen parameter: no device sync per call
        loss = 0 if self.training else None

        q16, qn = self._fuse(embedding)
        nranks = 1
        if self.world == 1:
            rec = self._retrieve(q16, qn, candidate_cells)
        else:
            import torch.distributed as dist

            pg, N = self.process_group, self.world
            if self.split_queries:
                # every rank holds B/N of the queries: all ranks need all fused queries + candidate lists
                Bl = q16.shape[0]
                cand_l = candidate_cells[:, :self.topk].to(torch.int64).contiguous()
                q_all = torch.empty((N * Bl, q16.shape[1]), dtype=q16.dtype, device=dev)
                qn_all = torch.empty((N * Bl,), dtype=qn.dtype, device=dev)
                cand_all = torch.empty((N * Bl, self.topk), dtype=torch.int64, device=dev)
                dist.all_gather_into_tensor(q_all, q16, group=pg)
                dist.all_gather_into_tensor(qn_all, qn, group=pg)
                dist.all_gather_into_tensor(cand_all, cand_l, group=pg)
                rec_all = self._retrieve(q_all, qn_all, cand_all)  # (N*Bl*topk, 4), rank-major over query owners
                rec = torch.empty_like(rec_all)  # chunk i = rank i's records of MY queries
                dist.all_to_all_single(rec, rec_all, group=pg)
                rec = rec.view(N, Bl * self.topk, 4)
            else:
                local = self._retrieve(q16, qn, candidate_cells)
                rec = torch.empty((N,) + tuple(local.shape), dtype=local.dtype, device=dev)
                dist.all_gather_into_tensor(rec, local, group=pg)
            nranks = N
        out = ops.proto_refine(rec, nranks, candidate_cells, candidate_probs, initial_preds, self.topk, temperature,
                               float(self.max_refinement), want_debug=return_debug)
        preds_llh, preds_geocell, guess_index = out[:3]
        if self.report_changed:  # :231-233 (one device sync, as in the reference)
            perc_changed = (guess_index != 0).sum() / guess_index.size(0)
            print(f"Changed geocell predictions of {perc_changed * 100:.1f} % of guesses.")
        if return_debug:
            return loss, preds_llh, preds_geocell, guess_index, out[3], out[4]
        return loss, preds_llh, preds_geocell

    def load_state_dict(self, state_dict, *a, **k):
        r = super().load_state_dict(state_dict, *a, **k)
        self._temperature_host = float(self.temperature.detach().cpu())
        return r
