#!/usr/bin/env python
"""Same CLI and output as the reference's bin/merge_pk.py: concatenate the per-slice P files along kx."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from saclaymocks_b200 import fitsio_lite as fitsio              # noqa: E402


def main():
    t0 = time.time()
    parser = argparse.ArgumentParser()
    parser.add_argument("-inDir")
    parser.add_argument("-outDir")
    parser.add_argument("-N", type=int, default=64)
    parser.add_argument("-NX", type=int, default=256)
    parser.add_argument("-NY", type=int, default=256)
    parser.add_argument("-NZ", type=int, default=256)
    a = parser.parse_args()
    print("Merging {} Pk fits files".format(a.N))
    cube = (a.NX == a.NY and a.NX == a.NZ)
    stem = "/P{}".format(a.NX) if cube else "/P{}-{}-{}".format(a.NX, a.NY, a.NZ)
    files = [fitsio.FITS(a.inDir + stem + "_{}_{}.fits".format(i, a.N)) for i in range(a.N)]
    head = files[0][0].read_header()
    hdict = {"Dcell": head["Dcell"], "NX": head["NX"], "NY": head["NY"], "NZ": head["NZ"]}
    out = fitsio.FITS(a.outDir + stem + ".fits", "rw", clobber=True)
    for ext in ("Pln1", "Pln2", "Pln3", "P0"):
        out.write(np.concatenate([f[ext].read() for f in files], axis=0), header=hdict, extname=ext)
    out.close()
    print("Merged fits file written in {}".format(a.outDir))
    print("Took {}s".format(time.time() - t0))


if __name__ == "__main__":
    main()
