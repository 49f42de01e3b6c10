#!/usr/bin/env python
"""GPU merge_spectra: same CLI and output files as the reference's bin/merge_spectra.py (argparse :21-42, file
selection :124-132, output :385-406).  Pieces of one forest are merged by position on the global pixel grid (equivalent
to the concatenate + argsort(wavelength) of :282-300); the small-scale field and FGPA (:303-339) run batched in
libsmk.so for all forests of the file.

Extra option: -noise {mt19937,philox}.  `mt19937` (default) draws delta_s from np.random exactly like the reference
(seed + islice, one normal(size=nz) per forest in (healpix, THING_ID) order); `philox` draws it on the GPU."""
import argparse
import glob
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from saclaymocks_b200 import constant                               # noqa: E402
from saclaymocks_b200 import fitsio_lite as fitsio                  # noqa: E402
from saclaymocks_b200.util import str2bool, radec2pix               # noqa: E402


def main():
    t_init = time.time()
    p = argparse.ArgumentParser()
    p.add_argument("-inDir")
    p.add_argument("-outDir")
    p.add_argument("-i", type=int)
    p.add_argument("-aa", type=float, default=-1)
    p.add_argument("-bb", type=float, default=1.58)
    p.add_argument("-cc", type=float, default=-1)
    p.add_argument("-paramfile", default=None)
    p.add_argument("-p1dfile", default=None)
    p.add_argument("-pixsize", type=float, default=0.2)
    p.add_argument("-nside", type=int, default=16)
    p.add_argument("-nest", default="True")
    p.add_argument("-rsd", default="True")
    p.add_argument("-addnoise", default="True")
    p.add_argument("-dla", default="False")
    p.add_argument("-zfix", type=float, default=None)
    p.add_argument("--fit-p1d", default="False")
    p.add_argument("--store-g", default="False")
    p.add_argument("-seed", type=int, default=None)
    p.add_argument("--check-id", default="True")
    p.add_argument("-ncpu", type=int, default=2)
    p.add_argument("-noise", choices=("mt19937", "philox"), default="mt19937")
    args = p.parse_args()
    import torch
    from saclaymocks_b200 import spectra as sp

    islice = args.i
    rsd, add_noise, dla = str2bool(args.rsd), str2bool(args.addnoise), str2bool(args.dla)
    store_g, check_id, nest = str2bool(args.store_g), str2bool(args.check_id), str2bool(args.nest)
    if str2bool(args.fit_p1d):
        print("--fit-p1d (offline tuning mode) is outside the hot path and is not supported here")
        sys.exit(1)
    seed = args.seed
    if seed is None:
        seed = int(np.random.randint(2 ** 31 - 1, size=1)[0])
        print("Seed has not been specified. Seed is set to {}".format(seed))
    else:
        seed = seed + islice
        print("Specified seed is {}".format(seed))
    np.random.seed(seed)
    qso_ids = None
    if check_id:
        ids = [fitsio.read(f, ext=1)["THING_ID"] for f in glob.glob(args.inDir + "/../qso/*")]
        qso_ids = np.concatenate(ids) if ids else np.zeros(0, dtype=np.int64)

    files = []
    for f in sorted(os.listdir(args.inDir)):                          # merge_spectra.py:124-132
        if f[f.rfind("-") + 1:f.find(".")] == str(islice):
            files.append(fitsio.FITS(args.inDir + "/" + f))
            print("{} opened".format(f))
    if not files:
        print("No fits file opened. Exit.")
        sys.exit()
    meta, header = [], None
    pieces = []
    for f in files:
        d = f[1].read()
        if len(d) == 0:
            continue
        if header is None:
            header = f[1].read_header()
        wav, delta = f["LAMBDA"].read(), f["DELTA_L"].read()
        eta = f["ETA_PAR"].read() if rsd else None
        velo = f["VELO_PAR"].read() if (rsd and dla) else None
        meta.append(d)
        pieces.append((wav, delta, eta, velo))
    meta_all = np.concatenate(meta)
    cpt1 = len(meta_all)
    npixeltot = header["Npixel"]
    z0, NX, dmax, ra0, dec0 = header["z0"], header["NX"], header["dmax"], header["ra0"], header["dec0"]
    print("IDs read - {} s".format(time.time() - t_init))

    # ---- geometry of the pixel grid: only the wavelength grid is needed; rebuild it from the header's cuts
    geom = sp.SkewerGeometry(32, 32, 32, 1.0, pixel=header["pixel"])
    geom_ok = geom.npixeltot == npixeltot
    if not geom_ok:
        print("pixel grid of the spectra files ({} pixels) differs from the default zmin/zmax grid ({}): "
              "stopping".format(npixeltot, geom.npixeltot))
        sys.exit(1)
    lam32 = np.float32(geom.lambda_vec)
    # ---- merge pieces by position (rows keyed by THING_ID)
    uid, first = np.unique(meta_all["THING_ID"], return_index=True)
    row_of = {int(i): r for r, i in enumerate(uid)}
    nq = len(uid)
    DL = np.full((nq, npixeltot), np.nan, dtype=np.float32)
    EP = np.zeros((nq, npixeltot), dtype=np.float32)
    VP = np.zeros((nq, npixeltot), dtype=np.float32)
    npieces = np.zeros(nq, dtype=int)
    for d, (wav, delta, eta, velo) in zip(meta, pieces):
        for r in range(len(d)):
            m = wav[r] > 0
            pos = np.searchsorted(lam32, wav[r][m])
            row = row_of[int(d["THING_ID"][r])]
            DL[row, pos] = delta[r][m]
            if eta is not None:
                EP[row, pos] = eta[r][m]
            if velo is not None:
                VP[row, pos] = velo[r][m]
            npieces[row] += 1
    M = meta_all[first]
    complete = ~np.isnan(DL).any(axis=1)
    if qso_ids is not None:
        known = np.isin(uid, qso_ids)
        for i in uid[~known]:
            print("WARNING ID: {} didn't match any QSO ID".format(i))
    else:
        known = np.ones(nq, dtype=bool)
    healpix = radec2pix(args.nside, M["RA"], M["DEC"], nest=nest)
    dev = torch.device("cuda:0")
    fg = sp.FGPA(geom, zfix=args.zfix, aa=args.aa, bb=args.bb, cc=args.cc, pixsize=args.pixsize,
                 p1dfile=args.p1dfile, paramfile=args.paramfile, device=dev)
    nfor = fg.forest_count(M["Z"])
    # ---- small scales: the reference draws one normal(size=nz) per known forest with a non-empty forest region, in
    #      (healpix, THING_ID) order, including forests it later drops for being incomplete (merge_spectra.py:246-350)
    order = np.lexsort((uid, healpix))
    t2 = time.time()
    DS = torch.zeros((nq, npixeltot), dtype=torch.float32, device=dev)
    if add_noise:
        if args.noise == "mt19937":
            noise = np.zeros((nq, fg.nfft_for(npixeltot)), dtype=np.float32)
            for r in order:
                if not known[r]:
                    continue
                n_row = int((~np.isnan(DL[r])).sum())
                if complete[r]:
                    if nfor[r] > 0:
                        noise[r] = np.random.normal(size=fg.nfft_for(n_row))
                else:                                   # incomplete row: the reference still consumes a draw
                    lam_rf = lam32[~np.isnan(DL[r])] / (np.float32(1) + np.float32(M["Z"][r]))
                    if ((lam_rf < np.float32(constant.lya)) & (lam_rf > np.float32(constant.lylimit))).any():
                        np.random.normal(size=fg.nfft_for(n_row))
            DS = fg.small_scales(np.where(complete, nfor, 0), noise=noise)
        else:
            DS = fg.small_scales(np.where(complete, nfor, 0), seed=seed, qso_ids=uid)
    dl_dev = torch.as_tensor(np.nan_to_num(DL, nan=-1e6), device=dev)
    F = fg.flux(dl_dev, DS if add_noise else None, torch.as_tensor(EP, device=dev) if rsd else None).cpu().numpy()
    DS = DS.cpu().numpy()
    print("FFT timer = {} s".format(time.time() - t2))

    names = ["RA", "DEC", "Z_noRSD", "Z", "HDU", "THING_ID", "PLATE", "MJD", "FIBERID", "PMF"]
    hlist = [{"name": "z0", "value": z0, "comment": "redshift of box center"}, {"name": "NX", "value": NX},
             {"name": "dmax", "value": dmax}, {"name": "ra0", "value": ra0, "comment": "right ascension of box center"},
             {"name": "dec0", "value": dec0, "comment": "declination of box center"}]
    cpt2 = cpt3 = 0
    good = complete & known
    for r in np.where(~complete & known)[0]:
        print("WARNING Spectrum hasn't the nominal lenght: ID {} has {} of {} pixels".format(
            uid[r], int((~np.isnan(DL[r])).sum()), npixeltot))
    for pix in np.unique(healpix):
        rows = np.where((healpix == pix) & good)[0]
        if len(rows) == 0:
            continue
        out = fitsio.FITS(args.outDir + "/spectra_merged-{}-{}.fits.gz".format(pix, islice), "rw", clobber=True)
        m = M[rows]
        out.write([m["RA"], m["DEC"], m["Z_noRSD"], m["Z"], np.full(len(rows), islice), uid[rows], m["PLATE"], m["MJD"],
                   m["FIBERID"], m["PMF"]], names=names, header=hlist, extname="METADATA")
        out.write(lam32, extname="LAMBDA")
        out.write(F[rows], extname="FLUX")
        if dla or store_g:
            out.write(DL[rows], extname="DELTA_L")
            out.write(np.float32(fg.growthf), extname="GROWTHF")
            out.write(VP[rows], extname="VELO_PAR")
        if store_g:
            out.write(EP[rows], extname="ETA_PAR")
            out.write(np.float32(fg.z), extname="Z")
            out.write(DS[rows], extname="DELTA_S")
        out.close()
        cpt3 += len(rows)
        cpt2 += int((npieces[rows] > 1).sum())
    print("Spectra merged and fits file saved.")
    print("{} initial forests.".format(cpt1))
    print("{} forest mergers.".format(cpt2))
    print("{} total forest written.".format(cpt3))
    print("Slice {} done. Took {}s".format(islice, time.time() - t_init))


if __name__ == "__main__":
    main()
