#!/usr/bin/env python
"""GPU draw_qso: same CLI, input files (boxln_{1,2,3}-<i>.fits, v{x,y,z}-<i>.fits) and output table
(QSO-<i>-<Nslice>.fits, bin/draw_qso.py:523-563) as the reference's bin/draw_qso.py; the cell loop runs in libsmk.so
(smk_draw_qso) on a B200.

Extra option: -draws {mt19937,philox}.  `mt19937` (default) generates the reference's legacy NumPy stream on the host
(np.random.seed(seed + i), five draws per z plane, draw_qso.py:154, 425-445) so that a given seed reproduces the
reference's catalogue; `philox` draws on the GPU, keyed by (seed, global cell index).
Not supported (the script stops): -random True (random catalogues), -desi True (needs etc/desi-healpix-weights.fits,
which the reference does not distribute), -zfix."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from saclaymocks_b200 import fitsio_lite as fitsio                 # noqa: E402
from saclaymocks_b200.util import str2bool                         # noqa: E402


def main():
    t_init = time.time()
    parser = argparse.ArgumentParser()
    parser.add_argument("-dmax", type=int, help="no quasar < dmax cells from edges, default 3", default=3)
    parser.add_argument("-indir", help="directory which contains boxes")
    parser.add_argument("-outpath", help="QSO fit path")
    parser.add_argument("-i", type=int, help="index of treated HDU")
    parser.add_argument("-Nslice", type=int, help="total number of slice")
    parser.add_argument("-chunk", type=int, help="index of treated chunk, default 0", default=0)
    parser.add_argument("-ra0", type=float, help="center box RA in degrees, default 0", default=0)
    parser.add_argument("-dec0", type=float, help="center box DEC in degrees, default 0", default=0)
    parser.add_argument("-dra", type=float, help="|ra-ra0|<dra in degrees, default -1 no cut", default=-1)
    parser.add_argument("-ddec", type=float, help="|dec-dec0|<ddec in degrees, default -1 no cut", default=-1)
    parser.add_argument("-random", type=str, help="If True, generate randoms. Default: False", default="False")
    parser.add_argument("-zmin", type=float, help="minimal redshift for drawing QSO, default is None", default=-1)
    parser.add_argument("-zmax", type=float, help="maximal redshift for drawing QSO, default is None", default=-1)
    parser.add_argument("-desi", type=str, help="select only objects in desi footprint", default="False")
    parser.add_argument("-zfix", type=float, help="not supported", default=None)
    parser.add_argument("-seed", type=int, help="specify a seed", default=None)
    parser.add_argument("-rsd", help="If True, rsd are added, default True", default="True")
    parser.add_argument("-dgrowthfile", help="dD/dz file, default etc/dgrowth.fits", default=None)
    parser.add_argument("-draws", choices=("mt19937", "philox"), default="mt19937")
    args = parser.parse_args()
    if str2bool(args.random) or str2bool(args.desi) or args.zfix is not None:
        print("draw_qso (GPU): -random True, -desi True and -zfix are not supported")
        sys.exit(1)
    rsd = str2bool(args.rsd)
    i_slice, Nslice = args.i, args.Nslice
    if args.seed is None:                                              # draw_qso.py:149-152
        seed = int(np.random.randint(2 ** 31 - 1, size=1)[0])
        print("Seed has not been specified. Seed is set to {}".format(seed))
    else:
        seed = args.seed + i_slice                                     # draw_qso.py:154
        print("Specified seed is {}".format(seed))

    import torch
    from saclaymocks_b200 import qso
    from saclaymocks_b200.boxes import BoxSynth
    dev = torch.device("cuda:0")
    head = fitsio.read_header(args.indir + "/boxln_1-{}.fits".format(i_slice))
    DX, NZ, NY, NXs, NX_full = head["DX"], head["NAXIS1"], head["NAXIS2"], head["NAXIS3"], head["NX"]
    print("Treating: slice = {} ; Nslice = {}\n".format(i_slice, Nslice))
    t0 = time.time()
    names = ["boxln_1", "boxln_2", "boxln_3"] + (["vx", "vy", "vz"] if rsd else [])
    box = {n: torch.from_numpy(np.ascontiguousarray(fitsio.read(args.indir + "/{}-{}.fits".format(n, i_slice)),
                                                    dtype=np.float32)).to(dev) for n in names}
    print("read boxes in {} s, shape: {}".format(time.time() - t0, tuple(box["boxln_1"].shape)))
    sigma_p = tuple(float(np.float32(qso.box_sigma(box[n]))) for n in names[:3])
    print("sigma(rho)=", sigma_p)
    st = qso.QsoSetup(NXs, NY, NZ, NX_full, DX, i_slice, Nslice, args.ra0, args.dec0, args.dra, args.ddec, args.zmin,
                      args.zmax, sigma_p, dmax=args.dmax)
    print("zmin, zmax:", st.z_min, st.z_max)
    print("norm={}".format(st.norm))
    bs = BoxSynth(16, 16, 24, DX, device=dev)                        # context / stream holder
    uni, rs = (None, None)
    if args.draws == "mt19937":
        uni, rs = qso.legacy_uniforms(seed, NXs, NY, NZ)
    t4 = time.time()
    cat = qso.QsoDrawer(bs).draw(st, [box[n] for n in names[:3]], [box[n] for n in names[3:]] if rsd else None,
                                 ix0=i_slice * NXs, uniforms=uni, seed=seed, chunk=args.chunk, rs=rs)
    print("End of loop on QSO. Took {} s".format(time.time() - t4))
    out_file = args.outpath + "/QSO-{}-{}.fits".format(i_slice, Nslice)
    qso.write_qso_file(out_file, cat, seed, args.ra0, args.dec0)
    print(len(cat["RA"]), "QSOs drawn")
    print(cat["nn_cond1"], "QSOs in the full box")
    print("Took {}s".format(time.time() - t_init))


if __name__ == "__main__":
    main()
