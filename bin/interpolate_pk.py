#!/usr/bin/env python
"""Same CLI and output file as the reference's bin/interpolate_pk.py (:31-47, 79-150): writes the kx rows
[i*NX/N, (i+1)*NX/N) of sqrt(P(|k|)/Vcell) for Pln1, Pln2, Pln3, P0 into P<...>_<i>_<N>.fits."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from saclaymocks_b200 import constant, pk                       # noqa: E402
from saclaymocks_b200 import fitsio_lite as fitsio              # noqa: E402
from saclaymocks_b200.cosmo import fgrowth                      # noqa: E402


def main():
    t0 = time.time()
    parser = argparse.ArgumentParser()
    parser.add_argument("-NX", type=int, default=256)
    parser.add_argument("-NY", type=int, default=256)
    parser.add_argument("-NZ", type=int, default=256)
    parser.add_argument("-pixel", type=float, default=2.19)
    parser.add_argument("-i", type=int, default=0)
    parser.add_argument("-N", type=int, default=64)
    parser.add_argument("-outDir")
    a = parser.parse_args()
    print("InterpolatePk : {}th slice over {}.".format(a.i, a.N))
    x0, x1 = a.i * a.NX // a.N, (a.i + 1) * a.NX // a.N
    if (x1 - x0) * a.NY * (a.NZ // 2 + 1) > 2 ** 31:
        print("Number of modes too large : > 2**31 !\nExit.")
        sys.exit(1)
    W = pk.weight_tables(a.NX, a.NY, a.NZ, a.pixel, x0, x1)
    name = ("/P{}_{}_{}.fits".format(a.NX, a.i, a.N) if (a.NY == a.NX and a.NZ == a.NX) else
            "/P{}-{}-{}_{}_{}.fits".format(a.NX, a.NY, a.NZ, a.i, a.N))
    f = fitsio.FITS(a.outDir + name, "rw", clobber=True)
    hdict = {"Dcell": a.pixel, "NX": a.NX, "NY": a.NY, "NZ": a.NZ}
    for j, z in enumerate((constant.z_QSO_bias_1, constant.z_QSO_bias_2, constant.z_QSO_bias_3)):
        f.write(W["Pln%d" % (j + 1)], header=hdict, extname="Pln%d" % (j + 1))
        f[-1].write_key("QSO_bias", pk.bias_qso(z), comment="QSO bias at z={}".format(z))
        f[-1].write_key("G", fgrowth(z, constant.omega_M_0), comment="growth factor at z={}".format(z))
        f[-1].write_key("Om", constant.omega_M_0, comment="Omega matter today")
    f.write(W["P0"], header=hdict, extname="P0")
    f.close()
    print("produced ", a.outDir + name)
    print("Took {}s".format(time.time() - t0))


if __name__ == "__main__":
    main()
