#!/usr/bin/env python
"""GPU make_spectra: same CLI and output files as the reference's bin/make_spectra.py (argparse :146-161, slab
loader :196-295, QSO-file selection :324-382, spectra files :528-597).  The per-quasar Python loop (:412-522) is one
batched smk_skewers launch over all quasars of the slab.

Multi-GPU: the reference runs one process per slice (-i); launched with torchrun and WITHOUT -i, rank r works through the
slices r, r + ranks, ... on its own GPU (cuda:LOCAL_RANK) and writes the same spectra-<slice>-<hdu>.fits.gz files."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from saclaymocks_b200 import constant                               # noqa: E402
from saclaymocks_b200 import fitsio_lite as fitsio                  # noqa: E402
from saclaymocks_b200.util import str2bool                          # noqa: E402


def main():
    t_init = time.time()
    parser = argparse.ArgumentParser()
    parser.add_argument("-dmax", type=int, default=3)
    parser.add_argument("-pixel", type=float, default=0.2)
    parser.add_argument("-zmin", type=float, default=1.3)
    parser.add_argument("-zmax", type=float, default=3.6)
    parser.add_argument("-QSOfile")
    parser.add_argument("-boxdir")
    parser.add_argument("-outDir")
    parser.add_argument("-i", type=int)
    parser.add_argument("-N", type=int)
    parser.add_argument("-NQSOfile", type=int, default=-1)
    parser.add_argument("-NQSO", type=int, default=-1)
    parser.add_argument("-rsd", default="True")
    parser.add_argument("-dla", default="False")
    parser.add_argument("-dgrowthfile", default=None)
    args = parser.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    slices = [args.i] if args.i is not None else list(range(rank, args.N, world))
    for iSlice in slices:
        run_slice(args, iSlice, t_init)


def run_slice(args, iSlice, t_init):
    import torch
    from saclaymocks_b200 import spectra as sp

    NSlice, dmax = args.N, args.dmax
    rsd, dla = str2bool(args.rsd), str2bool(args.dla)
    boxdir = args.boxdir
    print("Begining of MakeSpectra - {}".format(iSlice))
    head = fitsio.read_header(boxdir + "/box-0.fits", ext=0)
    DX, NZ, NY, nHDU = head["DX"], head["NAXIS1"], head["NAXIS2"], head["NX"]
    if head["NAXIS3"] != 1:
        print("NX=", head["NAXIS3"], " != 1  => abort !")
        sys.exit(1)
    if iSlice >= NSlice:
        print("iSlice=", iSlice, ">= NSlice =", NSlice, "=> abort !")
        sys.exit(1)
    if nHDU % NSlice != 0:
        print(NSlice, " slices not divider of nHDU:", nHDU, "=> abort !")
        sys.exit(1)
    geom = sp.SkewerGeometry(nHDU, NY, NZ, DX, zmin=args.zmin, zmax=args.zmax, pixel=args.pixel, dmax=dmax)
    iXmin = max((iSlice * nHDU) // NSlice - dmax, 0)                     # make_spectra.py:220-221
    iXmax = min(((iSlice + 1) * nHDU) // NSlice + dmax, nHDU)
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    t0 = time.time()

    def load_planes(name):
        return torch.as_tensor(np.concatenate([fitsio.read(boxdir + "/{}-{}.fits".format(name, ix), ext=0)
                                               for ix in range(iXmin, iXmax)]), device=dev)

    fields = {"box": load_planes("box")}
    if rsd:
        for n in ("eta_xx", "eta_yy", "eta_zz", "eta_xy", "eta_xz", "eta_yz"):
            fields[n] = load_planes(n)
        if dla:                                                           # v?-<k>.fits hold NX/N planes each
            per = nHDU // NSlice
            k0, k1 = iXmin // per, (iXmax - 1) // per + 1
            for n in ("vx", "vy", "vz"):
                v = np.concatenate([fitsio.read(boxdir + "/{}-{}.fits".format(n, k), ext=0) for k in range(k0, k1)])
                fields[n] = torch.as_tensor(np.ascontiguousarray(v[iXmin - k0 * per:iXmax - k0 * per]), device=dev)
    print("Boxes read. {} s".format(time.time() - t0))
    xSlicemin = geom.LX * iSlice / NSlice - geom.LX / 2
    xSlicemax = geom.LX * (iSlice + 1) / NSlice - geom.LX / 2
    print("Box {} - {} - {} with LX = {}, LY = {}, LZ = {}".format(nHDU, NY, NZ, geom.LX, geom.LY, geom.LZ))
    print("slice #", iSlice, "of box: ", xSlicemin, " < x < ", xSlicemax)
    npixeltot = geom.npixeltot
    # ---- QSO files: same conservative half selection as make_spectra.py:324-355
    NQSOfile = NSlice if args.NQSOfile < 0 else args.NQSOfile
    if iSlice >= NSlice // 2:
        ifile0, ifile1 = NSlice // 2, NQSOfile
        tanx_slice_max = (DX * (nHDU / NSlice) * (iSlice + 1) - geom.LX / 2) / (geom.R0 - geom.LZ / 2)
    else:
        ifile0, ifile1 = 0, NSlice // 2
        tanx_slice_max = np.abs((DX * (nHDU / NSlice) * iSlice - geom.LX / 2) / (geom.R0 - geom.LZ / 2))
    print("use QSO files:", ifile0, "to", ifile1 - 1)
    qsos, ra0, dec0 = [], None, None
    for ifile in range(ifile0, ifile1):
        name = args.QSOfile + str(ifile) + "-" + str(NQSOfile) + ".fits"
        try:
            f = fitsio.FITS(name, "r")
        except IOError:
            print("*Warning* Fits file {} cannot be read.".format(name))
            continue
        qsos.append(f[1].read())
        if ra0 is None:
            h = f[1].read_header()
            ra0, dec0 = h["RA0"], h["DEC0"]
    if not qsos or sum(len(q) for q in qsos) == 0:
        print("No QSO read. ==> Exit.")
        return
    qsos = np.concatenate(qsos)
    print(len(qsos), "QSO read")
    if args.NQSO > 0:
        qsos = qsos[:args.NQSO]
    zQSO = qsos["Z_QSO_RSD"] if rsd else qsos["Z_QSO_NO_RSD"]
    t0 = time.time()
    xyzr, nfor = sp.qso_lines_of_sight(geom, qsos["RA"], qsos["DEC"], zQSO, ra0, dec0)
    keep = (nfor >= 0) & (np.abs(xyzr[:, 0] / xyzr[:, 2]) <= tanx_slice_max)     # make_spectra.py:434-438
    idx = np.where(keep)[0]
    eng = sp.SkewerEngine(geom, device=dev)
    dl, ep, vp = (t.cpu().numpy() for t in eng.read_spec(fields, xyzr[idx], nfor[idx], ix0=iXmin, xmin=xSlicemin,
                                                          xmax=xSlicemax, rsd=rsd, dla=dla))
    owned = ~np.isnan(dl)                                                          # pixels of this slab
    has = owned.any(axis=1)                                                        # make_spectra.py:453-455
    idx, dl, ep, vp, owned = idx[has], dl[has], ep[has], vp[has], owned[has]
    if rsd and dla:
        vp = vp * geom.velo_rescale()[None, :]                                    # make_spectra.py:510
    print("End of loop: {}s".format(time.time() - t0))
    print("Writting...")
    t1 = time.time()
    lam32, z32 = np.float32(geom.lambda_vec), np.float32(geom.redshift)
    maxsize = int(owned.sum(axis=1).max()) if len(idx) else 0
    names = ["RA", "DEC", "Z_noRSD", "Z", "HDU", "THING_ID", "PLATE", "MJD", "FIBERID", "PMF"]
    hlist = [{"name": "z0", "value": constant.z0, "comment": "redshift of box center"},
             {"name": "pixel", "value": args.pixel}, {"name": "Npixel", "value": npixeltot},
             {"name": "NX", "value": 1}, {"name": "dmax", "value": dmax},
             {"name": "ra0", "value": ra0, "comment": "right ascension of box center"},
             {"name": "dec0", "value": dec0, "comment": "declination of box center"}]

    def padded(rows, fill):
        out = np.full((len(rows), maxsize), fill, dtype=np.float32)
        for r, v in enumerate(rows):
            out[r, :len(v)] = v
        return out

    q = qsos[idx]
    for ID in np.unique(q["HDU"]):
        m = np.where(q["HDU"] == ID)[0]
        out = fitsio.FITS(args.outDir + "/spectra-{}-{}.fits.gz".format(iSlice, ID), "rw", clobber=True)
        out.write([q["RA"][m], q["DEC"][m], q["Z_QSO_NO_RSD"][m], q["Z_QSO_RSD"][m], q["HDU"][m], q["THING_ID"][m],
                   q["PLATE"][m], q["MJD"][m], q["FIBERID"][m], q["PMF"][m]], names=names, header=hlist,
                  extname="METADATA")
        out.write(padded([lam32[owned[r]] for r in m], -1), extname="LAMBDA")
        out.write(padded([dl[r][owned[r]] for r in m], -2e6), extname="DELTA_L")
        if rsd:
            out.write(padded([ep[r][owned[r]] for r in m], -2e6), extname="ETA_PAR")
            if dla:
                out.write(padded([vp[r][owned[r]] for r in m], -2e6), extname="VELO_PAR")
        out.write(padded([z32[owned[r]] for r in m], -1), extname="REDSHIFT")
        out.close()
    print("Done. {} s".format(time.time() - t1))
    print(len(idx), "QSO written")
    print("Slice {} done. Took {}s".format(iSlice, time.time() - t_init))


if __name__ == "__main__":
    main()
