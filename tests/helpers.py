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
