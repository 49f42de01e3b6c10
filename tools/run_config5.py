#!/usr/bin/env python
"""BASELINE config 5: every chunk of the footprint (chunk_parameters(), bin/submit_mocks.py:611-674) through the
device-resident chain of one chunk -- boxes, quasars drawn on the resident boxes, sightlines, small-scale field, FGPA
(submit_mocks.py:375-425: run_boxes-<c>.sh -> run_chunk-<c>.sh) -- on the GPUs of one node, one process per GPU:

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_config5.py --box 2560 --out gpurun_out/config5.json

Per chunk: the rows are gathered to each quasar's home rank (ChunkPipeline.gather_rows) and rank 0 writes the DESI
transmission files of a sample of its quasars (saclaymocks_b200.transmissions.from_rows); the JSON holds the times of
every stage and checks on the outputs (window, redshift range, F in [0, 1], complete rows)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--box", type=int, default=2560)
    ap.add_argument("--out", default="gpurun_out/config5.json")
    ap.add_argument("--outdir", default="/tmp/smk_config5")
    ap.add_argument("--chunks", default="")
    ap.add_argument("--sample", type=int, default=200, help="quasars of rank 0 written as transmission files per chunk")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from saclaymocks_b200 import chunks, transmissions
    from saclaymocks_b200.chunk import ChunkPipeline
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nx = a.box
    ids = [int(c) for c in a.chunks.split(",")] if a.chunks else chunks.chunk_ids(nx)
    pipe = ChunkPipeline(nx, nx, 1536, 2.19, device=dev, rank=rank, nranks=world)
    t0 = time.time()
    pipe.set_weights({k: pipe.bs.weight_table(k) for k in ("Pln1", "Pln2", "Pln3", "P0")})
    torch.cuda.synchronize()
    t_weights = time.time() - t0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    res = []
    t_all = time.time()
    for c in ids:
        barrier()
        t0 = time.time()
        cat, out = pipe.run_chunk(chunk=c, seed=1000 + c)
        barrier()
        t_chunk = time.time() - t0
        t0 = time.time()
        rows = pipe.gather_rows()
        barrier()
        t_gather = time.time() - t0
        ra0, dra, dec0, ddec = chunks.chunk_window(nx, c)
        F = rows["flux"]
        n_home = int(F.shape[0])
        inside = ~torch.isnan(F)
        checks = {"window": bool(np.all(np.abs((cat["RA"] - ra0 + 180) % 360 - 180) < dra) and np.all(np.abs(cat["DEC"] - dec0) < ddec)),
                  "redshift": bool(np.all((cat["Z_QSO_RSD"] > 1.8) & (cat["Z_QSO_RSD"] < 3.6))),
                  "flux_range": bool(((F[inside] >= 0) & (F[inside] <= 1)).all()) if n_home else True,
                  "thing_id_unique": bool(len(np.unique(cat["THING_ID"])) == len(cat["THING_ID"]))}
        t0 = time.time()
        nfiles = 0
        if rank == 0 and n_home:
            k = min(a.sample, n_home)
            idx = rows["index"][:k]
            Fh = F[:k].cpu().numpy()
            lam = pipe.geom.lambda_vec
            os.makedirs(a.outdir + "/chunk%d" % c, exist_ok=True)
            nfiles = len(transmissions.from_rows(a.outdir + "/chunk%d" % c, cat["RA"][idx], cat["DEC"][idx],
                                                 cat["Z_QSO_NO_RSD"][idx], cat["Z_QSO_RSD"][idx], cat["THING_ID"][idx],
                                                 lam, np.nan_to_num(Fh, nan=1.0)))
        t_write = time.time() - t0
        tot = torch.tensor([n_home, int(inside.sum()), pipe.cat["own_pixels"]], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        res.append({"chunk": c, "window": [ra0, dra, dec0, ddec], "nqso": int(len(cat["RA"])), "home_rows": int(tot[0]),
                    "forest_pixels": int(tot[2]), "t_chunk_s": t_chunk, "phases_s_rank0": pipe.last_chunk_timings,
                    "t_gather_rows_s": t_gather,
                    "t_write_sample_s": t_write, "transmission_files_rank0": int(nfiles or 0),
                    "sigma_box": pipe.sigmas()["box"], "checks": checks})
        if rank == 0:
            print("chunk", c, json.dumps(res[-1]), flush=True)
    barrier()
    total = time.time() - t_all
    if rank == 0:
        cells = nx * nx * 1536
        line = {"config": "all %d chunks of the %d-cell layout through boxes + quasars + spectra" % (len(ids), nx),
                "n_gpus": world, "box": [nx, nx, 1536], "t_weight_tables_s": t_weights, "t_total_s": total,
                "chunk_cells_per_s": len(ids) * cells / sum(r["t_chunk_s"] for r in res), "chunks": res,
                "all_checks_ok": all(all(r["checks"].values()) for r in res)}
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        json.dump(line, open(a.out, "w"), indent=1)
        print(json.dumps({k: v for k, v in line.items() if k != "chunks"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
