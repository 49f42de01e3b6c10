"""world_size-2 gloo tests (CPU) of the host side of the multi-GPU path: the all-to-all buffer layouts around the
slab-decomposed 3-D transforms and the ownership rules of the skewer sharding."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from saclaymocks_b200 import slab
    NX, NY, NZ = 8, 12, 10
    P = NZ // 2 + 1 + 2                                   # padded rows like boxk
    rng = np.random.default_rng(0)
    box = rng.standard_normal((NX, NY, NZ)).astype(np.float32)
    ref = np.fft.rfftn(box)
    lo, hi = slab.x_planes(rank, world, NX)
    # forward: local z and y passes on the x-slab, exchange, x pass
    loc = np.zeros((hi - lo, NY, P), dtype=np.complex64)
    loc[:, :, :NZ // 2 + 1] = np.fft.fft(np.fft.rfft(box[lo:hi], axis=2), axis=1)
    send = torch.from_numpy(slab.forward_pack(loc, world))
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(-1).view(torch.float32), send.view(-1).view(torch.float32))
    boxk = np.fft.fft(slab.forward_unpack(recv.numpy(), world), axis=0)          # [NX][nyl][P]
    nyl = NY // world
    e1 = np.abs(boxk[:, :, :NZ // 2 + 1] - ref[:, rank * nyl:(rank + 1) * nyl]).max()
    # inverse: x pass, exchange, y pass, z pass -> the owned x-slab of the real box
    xk = np.fft.ifft(boxk, axis=0) * NX
    send = torch.from_numpy(np.ascontiguousarray(slab.inverse_pack(xk.astype(np.complex64), world)))
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(-1).view(torch.float32), send.view(-1).view(torch.float32))
    xs = slab.inverse_unpack(recv.numpy())                                         # [nxl][NY][P]
    back = np.fft.irfft(np.fft.ifft(xs[:, :, :NZ // 2 + 1], axis=1), n=NZ, axis=2) / NX
    e2 = np.abs(back - box[lo:hi]).max()
    # ownership: every forest pixel of every sightline is owned by exactly one rank, and by a rank that keeps it
    LX = 100.0
    xyzr = np.stack([rng.uniform(-40, 40, 50), rng.uniform(-40, 40, 50), rng.uniform(300, 400, 50),
                     rng.uniform(420, 500, 50)], 1)
    rvec = 200.0 + 0.2 * np.arange(1000)
    xmin, xmax = slab.x_bounds(rank, world, LX)
    keep = slab.touching(xyzr, rvec[0], rvec[-1], xmin, xmax)
    X = xyzr[:, 0:1] * rvec[None, :] / xyzr[:, 3:4]
    own = (X > xmin) & (X <= xmax)
    assert not own[~keep].any()                                                   # dropped quasars own nothing here
    t = torch.from_numpy(own.astype(np.int32))
    dist.all_reduce(t)
    e3 = int((t.numpy() != 1).sum())                                              # exactly one owner per pixel
    hl = slab.halo(rank, world, 3)
    assert hl == ((0, 3) if rank == 0 else (3, 0))
    assert (slab.home_rank(xyzr, world, LX) == (xyzr[:, 0] > 0)).all()
    with open(out + ".%d" % rank, "w") as f:
        f.write("%g %g %d" % (e1, e2, e3))
    dist.destroy_process_group()


def test_slab_exchange_layouts_gloo(tmp_path):
    out = str(tmp_path / "r")
    mp.spawn(_worker, args=(2, 29731, out), nprocs=2, join=True)
    for r in range(2):
        e1, e2, e3 = open(out + ".%d" % r).read().split()
        assert float(e1) < 1e-4 and float(e2) < 1e-5 and int(e3) == 0


def _gather_worker(rank, world, port, out):
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from saclaymocks_b200 import slab
    rng = np.random.default_rng(1)
    LX, n = 100.0, 60
    xyzr = np.stack([rng.uniform(-48, 48, n), rng.uniform(-40, 40, n), rng.uniform(300, 400, n),
                     rng.uniform(420, 500, n)], 1)
    rvec = 200.0 + 0.2 * np.arange(1200)
    X = xyzr[:, 0:1] * rvec[None, :] / xyzr[:, 3:4]
    truth = np.sin(X) + np.arange(n)[:, None]                       # what a single process would compute per pixel
    xmin, xmax = slab.x_bounds(rank, world, LX)
    mine = np.where(slab.touching(xyzr, rvec[0], rvec[-1], xmin, xmax))[0]
    rows = np.full((len(mine), len(rvec)), np.nan, dtype=np.float32)
    own = (X[mine] > xmin) & (X[mine] <= xmax)
    rows[own] = truth[mine][own]
    plan = slab.row_gather_plan(xyzr, rvec[0], rvec[-1], world, LX, rank)
    got = slab.exchange_rows(torch.from_numpy(rows), plan).numpy()
    home = plan["home_qso"]
    inside = (X[home] > -LX / 2) & (X[home] <= LX / 2)
    ok = np.array_equal(np.isnan(got), ~inside) and np.allclose(got[inside], np.float32(truth[home][inside]))
    cnt = torch.tensor([len(home)])
    dist.all_reduce(cnt)
    with open(out + ".%d" % rank, "w") as f:
        f.write("%d %d" % (int(ok), int(cnt.item())))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_rows_gather_to_home_rank_gloo(tmp_path, world):
    """Sharded spectra rows -> complete rows on the home rank (slab.row_gather_plan / exchange_rows): every pixel inside
    the box arrives exactly once, pixels outside stay NaN, every quasar has exactly one home."""
    out = str(tmp_path / "g")
    mp.spawn(_gather_worker, args=(world, 29741 + world, out), nprocs=world, join=True)
    for r in range(world):
        ok, cnt = open(out + ".%d" % r).read().split()
        assert int(ok) == 1 and int(cnt) == 60
