#!/usr/bin/env python
"""Train the geocell head on stored embeddings: the reference's `finetune_on_embeddings` (training/train_modes.py:
132-160) driven by its trainer loop (main_coordinator_idun_s3.py:383-424), on the B200 path end to end.

    python tools/train_head_from_sqlite.py data/sqlite/clip/dataset.sqlite --epochs 2 --batch 4096
    python tools/train_head_from_sqlite.py --synthetic 20000 --dim 1024 --dry-run       (no GPU: reads and reports)

Per batch, like the reference: labels (lng, lat) -> labels_clf = nearest centroid (:390-391; here the by-product of
gg_hav_row_stats, no second distance matrix), forward with the smoothed loss (:394), top-1 / top-5 accuracy
(:399-408, accumulated on the device: no .item() per batch), backward, AdamW.  Under torchrun the head is data
parallel (SuperGuessr.enable_data_parallel) and every rank reads its own stride of the table.
"""
import argparse
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geoguessr_ai_b200 import embedding_store as es  # noqa: E402


def synthetic_db(n, dim, seed=0):
    g = torch.Generator().manual_seed(seed)
    emb = torch.randn(n, 4, dim, generator=g)
    labels = torch.stack([torch.empty(n).uniform_(-180, 180, generator=g), torch.empty(n).uniform_(-60, 80, generator=g)], 1)
    path = os.path.join(tempfile.mkdtemp(prefix="gg_emb_"), "dataset.sqlite")
    es.write_embedding_sqlite(path, [f"loc{i:08d}" for i in range(n)], emb, labels)
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("sqlite", nargs="?", help="embedding SQLite (backend/s3bucket.py schema)")
    ap.add_argument("--synthetic", type=int, default=0, help="write and use a synthetic table of that many locations")
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--batch", type=int, default=4096, help="per GPU")
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--incomplete", default="drop", choices=["drop", "mean", "error"])
    ap.add_argument("--dry-run", action="store_true", help="read the table, print its shape, stop (no GPU needed)")
    args = ap.parse_args()

    path = synthetic_db(args.synthetic, args.dim) if args.synthetic else args.sqlite
    if not path:
        ap.error("give a SQLite path or --synthetic N")
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    t0 = time.perf_counter()
    table = es.read_embedding_sqlite(path, incomplete=args.incomplete, pin_memory=torch.cuda.is_available())
    N, V, D = table.embedding.shape
    if rank == 0:
        print(f"{path}: {N} locations x {V} headings x {D} dims ({table.embedding.numel() * 4 / 1e6:.1f} MB fp32), "
              f"{int((~table.present).sum())} filled slots, read in {time.perf_counter() - t0:.2f} s", flush=True)
    if args.dry_run:
        return
    if not torch.cuda.is_available():
        raise SystemExit("training needs an sm_100a GPU: the B200 path has no CPU fallback (use --dry-run to inspect a table)")

    import geoguessr_ai_b200 as gg
    from geoguessr_ai_b200 import ops

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        mine = torch.arange(rank, N, world)  # this rank's stride of the table
        table = es.EmbeddingTable([table.location_ids[i] for i in mine.tolist()], table.embedding[mine].pin_memory(),
                                  table.labels[mine].pin_memory(), table.present[mine], table.headings)
    model = gg.SuperGuessr(None, panorama=True, freeze_base=True, should_smooth_labels=True, embed_dim=D).to(dev).train()
    if world > 1:
        model.enable_data_parallel()
    opt = torch.optim.AdamW(model.cell_layer.parameters(), lr=args.lr, fused=True)
    table_xyz = ops.centroid_unit_vectors(model.geocell_centroid_coords.data)
    C = model.num_cells
    for epoch in range(args.epochs):
        stat = torch.zeros(4, device=dev)  # loss sum, top-1 hits, top-5 hits, samples
        t0 = time.perf_counter()
        for emb, labels, _ in es.iter_batches(table, args.batch, device=dev, shuffle=True, seed=epoch, drop_last=world > 1):
            _, labels_clf, _ = ops.hav_row_stats(labels, table_xyz, C, want_nearest=True)
            opt.zero_grad(set_to_none=True)
            out = model(embedding=emb, labels=labels, labels_clf=labels_clf)
            out.loss.backward()
            opt.step()
            with torch.no_grad():
                hit = out.top5_geocells.indices == labels_clf.unsqueeze(1)
                b = float(emb.shape[0])
                stat += torch.stack([out.loss.detach() * b, hit[:, 0].sum().float(), hit.any(1).sum().float(),
                                     torch.tensor(b, device=dev)])
        if world > 1:
            dist.all_reduce(stat)
        loss, top1, top5, n = stat.tolist()  # the epoch's only device -> host read
        dt = time.perf_counter() - t0
        if rank == 0:
            print(f"epoch {epoch}: loss {loss / n:.4f}  top-1 {top1 / n:.4f}  top-5 {top5 / n:.4f}  "
                  f"{n / dt / 1e6:.2f} M samples/s ({int(n)} samples, {dt:.2f} s)", flush=True)


if __name__ == "__main__":
    main()
