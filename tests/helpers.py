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


def philox4x32_10_np(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Random123) on uint64 arrays holding 32-bit words; checked against the known-answer
    vectors in tests/test_philox_cpu.py."""
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85, np.uint64(0xffffffff)
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & MASK for v in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xffffffff, int(k1) & 0xffffffff
    S = np.uint64(32)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> S) ^ c1 ^ np.uint64(k0)) & MASK, p1 & MASK, ((p0 >> S) ^ c3 ^ np.uint64(k1)) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & 0xffffffff, (k1 + W1) & 0xffffffff
    return c0, c1, c2, c3


def philox_normals_np(seed, ncells, cell0=0):
    """Host restatement of philox_normal4 (smk_philox.cuh): cells 4c..4c+3 from counter c = cell >> 2, Box-Muller."""
    assert cell0 % 4 == 0 and ncells % 4 == 0
    ctr = np.arange(cell0 // 4, (cell0 + ncells) // 4, dtype=np.uint64)
    c = philox4x32_10_np(ctr & np.uint64(0xffffffff), ctr >> np.uint64(32), 0 * ctr, 0 * ctr, seed & 0xffffffff, seed >> 32)
    S = 2.0 ** -24
    u0, u1 = ((c[0] >> np.uint64(8)).astype(np.float64) + 1) * S, (c[1] >> np.uint64(8)).astype(np.float64) * S
    u2, u3 = ((c[2] >> np.uint64(8)).astype(np.float64) + 1) * S, (c[3] >> np.uint64(8)).astype(np.float64) * S
    r0, r1 = np.sqrt(-2 * np.log(u0)), np.sqrt(-2 * np.log(u2))
    out = np.stack([r0 * np.cos(2 * np.pi * u1), r0 * np.sin(2 * np.pi * u1), r1 * np.cos(2 * np.pi * u3),
                    r1 * np.sin(2 * np.pi * u3)], axis=1)
    return out.reshape(-1)
