#!/usr/bin/env python
"""make_transmissions with the reference's CLI (bin/make_transmissions.py:22-31): regroups the spectra_merged files by
healpix pixel into DESI transmission-<nside>-<pix>.fits.gz files.  Pure I/O (no GPU); the output directories
<outDir>/<pix//100>/<pix>/ are created when missing (the reference relies on submit_mocks.py having made them)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from saclaymocks_b200 import fitsio_lite as fitsio                 # noqa: E402
from saclaymocks_b200 import transmissions                         # noqa: E402
from saclaymocks_b200.util import str2bool                         # noqa: E402


def main():
    t_init = time.time()
    parser = argparse.ArgumentParser()
    parser.add_argument("-inDir", help="spectra file, default DesiMocks/spectra/", default="DesiMocks/spectra/")
    parser.add_argument("-outDir", help="out directory. Default DesiMocks/output/", default="DesiMocks/output/")
    parser.add_argument("-job", type=int, help="index of current job", default=0)
    parser.add_argument("-ncpu", type=int, help="total number of cpu", default=64)
    parser.add_argument("-nside", type=int, help="nside for healpix. Default 16", default=16)
    parser.add_argument("-nest", help="If True, healpix scheme is nest. Default True", default="True")
    parser.add_argument("-prod", help="accepted for compatibility", default="True")
    parser.add_argument("-dla", help="If True, add the DLA HDU from <outDir>/master_DLA.fits, default False",
                        default="False")
    args = parser.parse_args()
    nside, job, ncpu = args.nside, args.job, args.ncpu
    npixel = 12 * nside * nside
    last = npixel if job == ncpu - 1 else int((job + 1) * npixel / ncpu)       # make_transmissions.py:44-47
    pixels = range(int(job * npixel / ncpu), last)
    print("Treated healpix pixels: {}".format(pixels))
    dla_cat = None
    if str2bool(args.dla):
        dla_cat = fitsio.read(args.outDir + "/master_DLA.fits", ext=1)
    t0 = time.time()
    cpt = transmissions.from_merged_files(args.inDir, args.outDir, pixels, nside=nside, nest=str2bool(args.nest),
                                          dla_cat=dla_cat)
    if cpt == 0:
        print("No file found for these healpix pixels")
        sys.exit()
    print("{} Sorted spectra written. {} s".format(cpt, time.time() - t0))
    print("Job {} done. Took {}s".format(job, time.time() - t_init))


if __name__ == "__main__":
    main()
