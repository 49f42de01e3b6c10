"""Multi-GPU parity: the x-slab / y-slab sharded pipeline (all-to-all transposes, halo exchange, skewers sharded by
owning slab) against the CPU oracle, at 2, 4 and 8 ranks, including the 1024-, 2048- and 2560-point fused-exchange x
passes of the benched configurations.  A case needs as many GPUs as ranks (skipped otherwise): `gpurun --gpus 8` runs
all of them; the log of that run is kept under profiles/."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, shape, dcell, out, p2p, xsms=0):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["SMK_P2P"] = p2p
    os.environ["SMK_X_SMS"] = str(xsms)          # fused-exchange x pass: one CTA per tile (0) or persistent on that many
                                                 # CTAs (the default of 8 ranks, 96, is what bench.py's self-check runs)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from oracle import boxes as ob
    from oracle import pk_weights
    from oracle import spectra as osp
    from saclaymocks_b200.chunk import ChunkPipeline
    from saclaymocks_b200 import spectra as sp
    from helpers import rel_l2
    NX, NY, NZ = shape
    W = pk_weights.weights(NX, NY, NZ, dcell)
    noise = ob.draw_noise(NX, NY, NZ, 42)
    raw, p0, boxes, sig = ob.make_boxes(NX, NY, NZ, dcell, 42, W, noise=noise)
    pipe = ChunkPipeline(NX, NY, NZ, dcell, device=dev, rank=rank, nranks=world, zfix=2.4)
    nyl, nxl = NY // world, NX // world
    pipe.set_weights({k: v[:, rank * nyl:(rank + 1) * nyl] for k, v in W.items()})
    # synthetic quasars inside the box
    rng = np.random.default_rng(5)
    geom = pipe.geom
    half = np.degrees(np.arctan((geom.LX / 2 - 3 * dcell) / (geom.R0 + geom.LZ / 2))) * 0.95
    nq = 24
    ra = (190.0 + rng.uniform(-half, half, nq)).astype("f4")
    dec = (0.0 + rng.uniform(-half, half, nq)).astype("f4")
    z = rng.uniform(2.0, 3.5, nq).astype("f4")
    pipe.set_catalogue(ra, dec, z, 190.0, 0.0)
    pipe.step_boxes(noise=torch.as_tensor(noise[rank * nxl:(rank + 1) * nxl].copy(), device=dev))
    errs = {}
    for name in ob.PRODUCTS:
        got = pipe.interior(name).cpu().numpy()
        errs[name] = rel_l2(got, boxes[name][rank * nxl:(rank + 1) * nxl])
    sg = pipe.sigmas()
    errs["sigma"] = max(abs(sg[n] / sig[n] - 1) for n in ob.PRODUCTS)
    # skewers: every rank computes the pixels its slab owns; compare with the oracle's per-slab pieces
    dl, ep, vp, F = (t.cpu().numpy() for t in pipe.step_skewers())
    og = osp.Geometry(NX, NY, NZ, dcell)
    q = np.zeros(nq, dtype=[("RA", "f4"), ("DEC", "f4"), ("Z_QSO_NO_RSD", "f4"), ("Z_QSO_RSD", "f4"),
                            ("THING_ID", "i8"), ("HDU", "i4")])
    q["RA"], q["DEC"], q["Z_QSO_RSD"], q["Z_QSO_NO_RSD"], q["THING_ID"] = ra, dec, z, z, np.arange(nq)
    files = [q] * world            # every "file" holds the whole catalogue: the half-selection then keeps all of it
    lam32 = np.float32(geom.lambda_vec)
    worst = 0.0
    npieces = 0
    sel = list(pipe.cat["sel"])
    for p in osp.make_spectra_slice(og, boxes, [q[:0]] * (world // 2) + [q] + [q[:0]] * (world - world // 2 - 1)
                                    if rank >= world // 2 else [q] + [q[:0]] * (world - 1), rank, world, 190.0, 0.0):
        idx = np.searchsorted(lam32, p["lam"])
        row = sel.index(p["id"])
        m = p["delta_l"] > -1e5
        worst = max(worst, float(np.max(np.abs(dl[row, idx][m] - p["delta_l"][m]))) if m.any() else 0.0,
                    float(np.max(np.abs(ep[row, idx] - p["eta_par"]))))
        assert np.all(dl[row, idx][~m] == -1e6)
        npieces += 1
    errs["skewers"] = worst
    errs["npieces"] = npieces
    # complete rows on the home rank (ChunkPipeline.gather_rows): every piece of every slice of this rank's quasars
    rows = pipe.gather_rows()
    gdl, gep = rows["delta_l"].cpu().numpy(), rows["eta_par"].cpu().numpy()
    home = list(rows["index"])
    worst_g, npx_g = 0.0, 0
    if home:
        qh = q[np.asarray(home)]
        for s_ in range(world):
            files_ = [qh[:0]] * world
            files_[min(s_, world - 1) if s_ < world // 2 else world // 2] = qh     # a file the half-selection keeps
            for p in osp.make_spectra_slice(og, boxes, files_, s_, world, 190.0, 0.0):
                idx = np.searchsorted(lam32, p["lam"])
                row = home.index(p["id"])
                m = p["delta_l"] > -1e5
                if m.any():
                    worst_g = max(worst_g, float(np.max(np.abs(gdl[row, idx][m] - p["delta_l"][m]))),
                                  float(np.max(np.abs(gep[row, idx] - p["eta_par"]))))
                    npx_g += int(m.sum())
    errs["gathered"] = worst_g
    errs["gathered_px"] = npx_g
    import json
    json.dump({k: float(v) for k, v in errs.items()}, open(out + ".%d" % rank, "w"))
    dist.destroy_process_group()


CASES = [(2, (32, 32, 96), "auto", 0), (2, (64, 32, 96), "0", 0),
                                                  (2, (2048, 32, 96), "1", 0), (4, (64, 64, 96), "auto", 0),
                                                  (2, (64, 32, 96), "1", 3), (2, (2048, 32, 96), "1", 5),
                                                  # the transform lengths and rank counts bench.py runs (configs 3, 4)
                                                  (2, (1024, 32, 96), "1", 0), (4, (1024, 32, 96), "1", 0),
                                                  (4, (2048, 32, 96), "1", 0), (4, (64, 64, 96), "0", 0),
                                                  (8, (64, 64, 96), "auto", 0), (8, (64, 64, 96), "0", 0),
                                                  (8, (2048, 64, 96), "1", 0), (8, (2560, 64, 96), "1", 0),
                                                  (8, (2560, 64, 96), "1", 24)]


@pytest.mark.parametrize("world,shape,p2p,xsms", CASES,
                         ids=["w%d-%dx%dx%d-p2p%s-xsms%d" % (w, s[0], s[1], s[2], p, x) for w, s, p, x in CASES])
def test_sharded_pipeline_matches_oracle(tmp_path, world, shape, p2p, xsms):
    """p2p: "1" = fused peer-store exchange (tiled receive layout), "0" = NCCL all-to-all, "auto" = by size;
    xsms > 0: the fused x pass runs persistent on that many CTAs (SMK_X_SMS), several tiles per CTA."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    out = str(tmp_path / "res")
    port = 29600 + 20 * CASES.index((world, shape, p2p, xsms))
    mp.spawn(_worker, args=(world, port, shape, 3364.0 / shape[2], out, p2p, xsms), nprocs=world, join=True)
    pieces = gathered = 0
    for r in range(world):
        import json
        errs = json.load(open(out + ".%d" % r))
        for k, v in errs.items():
            if k == "npieces":
                pieces += v
            elif k == "gathered_px":
                gathered += v
            elif k == "gathered":
                assert v < 1e-5, (r, k, v)
            elif k == "skewers":
                assert v < 1e-5, (r, k, v)
            elif k == "sigma":
                assert v < 1e-4, (r, k, v)
            else:
                assert v < 1e-5, (r, k, v)
    assert pieces > 0 and gathered > 1000
