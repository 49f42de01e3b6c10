"""Shared test helpers: rebuild the golden runs' inputs and drive the oracle the way the reference CLIs were driven."""
import numpy as np

QSO_DTYPE = [("RA", "f4"), ("DEC", "f4"), ("Z_QSO_NO_RSD", "f4"), ("Z_QSO_RSD", "f4"), ("THING_ID", "i8"), ("HDU", "i4")]


def qso_files_from_golden(g):
    """The synthetic QSO-<i>-<nslice>.fits tables of tests/golden/run_reference_shimmed.py, one array per file."""
    n = len(g["qso_RA"])
    q = np.zeros(n, dtype=QSO_DTYPE)
    for name, _ in QSO_DTYPE:
        q[name] = g["qso_" + name]
    nslice = int(g["nslice"])
    return [q[q["HDU"] == i] for i in range(nslice)]


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128 if np.iscomplexobj(a) else np.float64)
    b = np.asarray(b).astype(np.complex128 if np.iscomplexobj(b) else np.float64)
    return np.sqrt(np.sum(np.abs(a - b) ** 2) / np.sum(np.abs(b) ** 2))


def write_qso_files(g, qsodir):
    """QSO-<i>-<nslice>.fits tables (layout of bin/draw_qso.py:523-563) holding the golden run's synthetic catalogue."""
    from saclaymocks_b200 import fitsio_lite as fitsio
    nslice = int(g["nslice"])
    names = ["Z_QSO_NO_RSD", "Z_QSO_RSD", "RA", "DEC", "HDU", "THING_ID", "PLATE", "MJD", "FIBERID", "PMF",
             "XX", "YY", "ZZ"]
    for i, q in enumerate(qso_files_from_golden(g)):
        n = len(q)
        tid = q["THING_ID"]
        pmf = np.array(["{}-{}-{}".format(t, 0, 0) for t in tid], dtype="S21")
        cols = [q["Z_QSO_NO_RSD"], q["Z_QSO_RSD"], q["RA"], q["DEC"], q["HDU"], tid, tid.copy(), np.zeros(n, "i4"),
                np.zeros(n, "i4"), pmf, np.zeros(n, "f4"), np.zeros(n, "f4"), np.zeros(n, "f4")]
        f = fitsio.FITS(qsodir + "/QSO-{}-{}.fits".format(i, nslice), "rw", clobber=True)
        f.write(cols, names=names, header=[{"name": "seed", "value": int(g["seed"])},
                                          {"name": "ra0", "value": float(g["ra0"])},
                                          {"name": "dec0", "value": float(g["dec0"])}], extname="QSO")
        f.close()
