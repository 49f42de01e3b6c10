#!/usr/bin/env python
"""GPU make_boxes: same CLI, file names and FITS headers as the reference's bin/make_boxes.py
(argparse :137-150, products :242-431, FITS layout :110-117); the arithmetic runs in libsmk.so on a B200.

Multi-GPU: launched with torchrun (one process per GPU, e.g. `torchrun --nproc-per-node 8 bin/make_boxes.py -NX 2560
...`) the box is x-slab sharded over the ranks (saclaymocks_b200/chunk.py); every rank writes the files that hold its
own planes -- the same file names and headers as one process would write (box-<ix>.fits and eta_*-<ix>.fits hold one
plane each, boxln_<k>-<i>.fits and v?-<i>.fits NX/nHDU planes each, so nHDU must be a multiple of the rank count), with
sigma reduced over all ranks.  boxk.npy (the resume file of the reference) is written as one y-slab shard per rank,
boxk-<rank>of<n>.npy.  `-PkDir gpu` evaluates the spectral weight tables on the GPU from the P(k) splines
(smk_pk_weights, bit-equal to the P-file to > 99.9 %) instead of reading the 4 x NX*NY*(NZ/2+1) table file.

Extra option: -noise {philox,mt19937}.  `mt19937` draws the white noise on the host exactly like the reference
(np.random.seed(seed) + one np.random.normal plane per iz, make_boxes.py:46-48, 163-171) so that a given seed
reproduces the reference's boxes; `philox` (default) draws it inside the forward z pass on the GPU.
-ncpu is accepted and ignored (no FFTW wisdom on a GPU; nothing is written into $SACLAYMOCKS_BASE/etc)."""
import argparse
import glob
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from saclaymocks_b200 import constant, tables                      # noqa: E402
from saclaymocks_b200 import fitsio_lite as fitsio                 # noqa: E402
from saclaymocks_b200.util import str2bool                         # noqa: E402


def write_box(box, boxfilename, nHDU, Dcell, NX, NY, NZ, sigma, seed):
    """make_boxes.py:110-117."""
    for i in range(nHDU):
        f = fitsio.FITS(boxfilename + "-{}.fits".format(i), "rw", clobber=True)
        f.write(box[i * NX // nHDU:(i + 1) * NX // nHDU],
                header={"DX": Dcell, "DY": Dcell, "DZ": Dcell, "NX": NX, "NY": NY, "NZ": NZ})
        if i == 0:
            f[0].write_key("sigma", np.float32(sigma), comment="std of the box")
            f[0].write_key("seed", np.int32(seed), comment="seed used to generate randoms")
        f.close()


def nfiles(prefix):
    return len(glob.glob(prefix + "*"))


def main_sharded(args, seed, rsd, NX, NY, NZ, dd):
    """One rank of a torchrun launch: this rank's x-slab of every product, written to the files that hold its planes."""
    import torch
    import torch.distributed as dist
    from saclaymocks_b200.chunk import ChunkPipeline
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sd = torch.tensor([seed], dtype=torch.int64, device=dev)       # a seed drawn at random must be the same on every rank
    dist.broadcast(sd, 0)
    seed = int(sd.item())
    Dcell, nHDU, outDir = args.pixel, args.nHDU, args.outDir
    if NX % world or NY % world or nHDU % world:
        raise ValueError("NX, NY and nHDU must be multiples of the number of ranks ({})".format(world))
    t_init = time.time()
    pipe = ChunkPipeline(NX, NY, NZ, Dcell, device=dev, rank=rank, nranks=world, rsd=rsd)
    bs = pipe.bs
    bs.dgrowth0 = float(dd[0])
    nxl, nyl = NX // world, NY // world
    if args.PkDir == "gpu":
        W = {k: bs.weight_table(k) for k in ("Pln1", "Pln2", "Pln3", "P0")}
    else:
        Pfilename = args.PkDir + ("/P{}.fits".format(NX) if (NY == NX and NZ == NX) else
                                  "/P{}-{}-{}.fits".format(NX, NY, NZ))
        W = {k: np.ascontiguousarray(fitsio.read(Pfilename, ext=k)[:, rank * nyl:(rank + 1) * nyl])
             for k in ("Pln1", "Pln2", "Pln3", "P0")}
    pipe.set_weights(W)
    noise = None
    if args.noise == "mt19937":            # the reference's stream: every rank draws all of it and keeps its own planes
        np.random.seed(seed)
        slab_noise = np.zeros((nxl, NY, NZ), dtype=np.float32)
        for iz in range(NZ):
            slab_noise[:, :, iz] = np.float32(np.random.normal(size=[NX, NY]))[rank * nxl:(rank + 1) * nxl]
        noise = torch.as_tensor(slab_noise, device=dev)
    products = [n for n in bs_products() if rsd or n.startswith("box")]
    pipe.step_boxes(seed=seed, noise=noise, products=products)
    torch.cuda.synchronize()
    np.save(outDir + "/boxk-{}of{}.npy".format(rank, world), bs.boxk_to_numpy(pipe.boxk))
    if rank == 0:
        np.save(outDir + "/seed_boxk.npy", seed)
    sig = pipe.sigmas()                                                        # all-reduced over the ranks
    for name in products:
        if not (sig[name] > 0) or np.isnan(sig[name]):                         # make_boxes.py:100-105
            raise ValueError("box {} is null".format(name))
        box = pipe.interior(name).cpu().numpy()
        per_plane = name == "box" or name.startswith("eta_")
        nfile = NX if per_plane else nHDU
        per = NX // nfile                                                      # planes per file
        for i in range(rank * nxl // per, (rank + 1) * nxl // per):
            f = fitsio.FITS(outDir + "/" + name + "-{}.fits".format(i), "rw", clobber=True)
            f.write(box[i * per - rank * nxl:(i + 1) * per - rank * nxl],
                    header={"DX": Dcell, "DY": Dcell, "DZ": Dcell, "NX": NX, "NY": NY, "NZ": NZ})
            if i == 0:
                f[0].write_key("sigma", np.float32(sig[name]), comment="std of the box")
                f[0].write_key("seed", np.int32(seed), comment="seed used to generate randoms")
            f.close()
        if rank == 0:
            print(name, "sigma = {}".format(sig[name]))
    dist.barrier()
    if rank == 0:
        print("NX=", NX, "ranks=", world)
        print("Took {}s".format(time.time() - t_init))
    dist.destroy_process_group()


def bs_products():
    from saclaymocks_b200.boxes import PRODUCTS
    return PRODUCTS


def main():
    print("Starting DrawGRF...")
    t_init = time.time()
    parser = argparse.ArgumentParser()
    parser.add_argument("-pixel", type=float, help="pixel size (Mpc/h), default 2.19", default=2.19)
    parser.add_argument("-NX", type=int, help="number of pixels along x, default 256", default=256)
    parser.add_argument("-NY", type=int, help="number of pixels along y, default = NX", default=-1)
    parser.add_argument("-NZ", type=int, help="number of pixels along z, default = NX", default=-1)
    parser.add_argument("-nHDU", type=int, help="number of HDU box.fits, default 1", default=1)
    parser.add_argument("-ncpu", type=int, default=2)
    parser.add_argument("-PkDir", help="directory of Pk fits file")
    parser.add_argument("-seed", type=int, help="specify a seed", default=None)
    parser.add_argument("-rsd", type=str, help="If True, rsd are added, default True", default="True")
    parser.add_argument("-dgrowthfile", help="dD/dz file, default etc/dgrowth.fits", default=None)
    parser.add_argument("-outDir", help="directory where the box are saved")
    parser.add_argument("-noise", choices=("philox", "mt19937"), default="philox")
    args = parser.parse_args()
    import torch
    from saclaymocks_b200.boxes import BoxSynth, WEIGHT_OF

    rsd = str2bool(args.rsd)
    Dcell, NX = args.pixel, args.NX
    NY = NX if args.NY < 0 else args.NY
    NZ = NX if args.NZ < 0 else args.NZ
    seed = args.seed
    if seed is None:
        seed = int(np.random.randint(2 ** 31 - 1, size=1)[0])
        print("Seed has not been specified. Seed is set to {}".format(seed))
    else:
        print("Specified seed is {}".format(seed))
    nHDU, outDir = args.nHDU, args.outDir
    Pfilename = args.PkDir + ("/P{}.fits".format(NX) if (NY == NX and NZ == NX) else
                              "/P{}-{}-{}.fits".format(NX, NY, NZ))          # make_boxes.py:186-189
    z, dd, om = tables.dgrowth(args.dgrowthfile)
    if rsd and constant.omega_M_0 != om:                                       # make_boxes.py:309-311
        raise ValueError("Omega_M_0 in constant ({}) != OM in dgrowth file ({})".format(constant.omega_M_0, om))

    if int(os.environ.get("WORLD_SIZE", "1")) > 1:                            # torchrun: x-slab sharded over the ranks
        return main_sharded(args, seed, rsd, NX, NY, NZ, dd)
    dev = torch.device("cuda:0")
    bs = BoxSynth(NX, NY, NZ, Dcell, device=dev)
    bs.dgrowth0 = float(dd[0])
    boxkfile = outDir + "/boxk.npy"
    boxk_exist = os.path.isfile(boxkfile)
    t0 = time.time()
    if boxk_exist:                                                             # resume, make_boxes.py:209-228
        print("{} already exists ! Reading boxk.npy file to compute density and velocity boxes...".format(boxkfile))
        a = np.load(boxkfile)
        try:
            seed = int(np.load(outDir + "/seed_boxk.npy"))
        except Exception:
            print("WARNING: didn't find {}/seed_boxk.npy".format(outDir))
        sigma_k = a.std()
        if sigma_k > 70 * NX:          # boxk.npy already holds boxk*P0 (make_boxes.py:218-226): undo and save, then go on
            print("Sigma of boxk is {} > 70*{}:".format(sigma_k, NX))
            print("dividing boxk by P0(k) and saving...")
            with np.errstate(divide="ignore", invalid="ignore"):
                a /= fitsio.read(Pfilename, ext="P0")
            a[0, 0, 0] = 0j
            np.save(boxkfile, a)
        else:
            print("Sigma of boxk is {} < 70*{}".format(sigma_k, NX))
        boxk = bs.boxk_from_numpy(a)
        del a
    else:
        print(NX, NY, NZ)
        if args.noise == "mt19937":
            np.random.seed(seed)
            noise = np.zeros((NX, NY, NZ), dtype=np.float32)
            for iz in range(NZ):
                noise[:, :, iz] = np.float32(np.random.normal(size=[NX, NY]))
            boxk = bs.draw_grf_boxk(noise=torch.as_tensor(noise, device=dev))
            del noise
        else:
            boxk = bs.draw_grf_boxk(seed=seed)
        k2 = float((boxk.abs() ** 2).sum())
        if not (k2 > 0) or np.isnan(k2):                                       # make_boxes.py:63-70
            raise ValueError("boxk is null")
        np.save(boxkfile, bs.boxk_to_numpy(boxk))
        np.save(outDir + "/seed_boxk.npy", seed)
        print("boxk produced and saved:", time.time() - t0, " s ")

    def product(name, nfile, wname=None):
        boxfile = outDir + "/" + name
        pattern = boxfile + "-" if name == "box" else boxfile
        if nfile == nfiles(pattern) and boxk_exist and name != "box":
            print("{} files already exist ! Skiping this step.".format(boxfile))
            return
        t1 = time.time()
        wt = bs.upload_weights(fitsio.read(Pfilename, ext=wname)) if wname else None
        box, stats = bs.synth(boxk, name, wtable=wt, store_p0=True)
        sigma = bs.sigma(stats)                                                # raises ValueError on a null box
        print("FFT done", time.time() - t1, "s")
        print("sigma = {}".format(sigma))
        write_box(box.cpu().numpy(), boxfile, nfile, Dcell, NX, NY, NZ, sigma, seed)
        print(boxfile, "written", time.time() - t1, "s")

    print("Computing delta boxes...")
    for i in (1, 2, 3):
        product("boxln_%d" % i, nHDU, "Pln%d" % i)
    product("box", NX, "P0")                                                   # boxk <- boxk*P0 (make_boxes.py:289-291)
    np.save(boxkfile, bs.boxk_to_numpy(boxk))
    if rsd:
        print("Computing eta boxes:")
        for name in ("eta_xx", "eta_yy", "eta_zz", "eta_xy", "eta_xz", "eta_yz"):
            print(name + "...")
            product(name, NX)
        print("Computing velocity boxes:")
        for name in ("vx", "vy", "vz"):
            print(name + "...")
            product(name, nHDU)
    print("NX=", NX, "nCPU=", args.ncpu)
    print("Took {}s".format(time.time() - t_init))


if __name__ == "__main__":
    main()
