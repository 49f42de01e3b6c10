"""Static calibration tables of the reference's etc/ directory.

Lookup order: $SACLAYMOCKS_BASE/etc/<file>.fits when that directory exists (drop-in for the
reference installation, bin/make_boxes.py:308, bin/merge_spectra.py:59,108), else the packaged
copy saclaymocks_b200/data/etc_tables.npz (made by tests/golden/make_etc_tables.py)."""
import os

import numpy as np

from . import fitsio_lite as fitsio

_NPZ = os.path.join(os.path.dirname(__file__), "data", "etc_tables.npz")
_cache = {}


def _etc_dir():
    base = os.environ.get("SACLAYMOCKS_BASE")
    if base and os.path.isdir(os.path.join(base, "etc")):
        return os.path.join(base, "etc")
    return None


def _npz():
    if "npz" not in _cache:
        _cache["npz"] = dict(np.load(_NPZ))
    return _cache["npz"]


def params(filename=None):
    """z, a, b, c of etc/params.fits (bin/merge_spectra.py:58-75)."""
    etc = _etc_dir()
    if filename is None and etc and os.path.isfile(etc + "/params.fits"):
        filename = etc + "/params.fits"
    if filename is not None:
        d = fitsio.read(filename, ext=1)
        return d["z"], d["a"], d["b"], d["c"]
    t = _npz()
    return t["params_z"], t["params_a"], t["params_b"], t["params_c"]


def dgrowth(filename=None):
    """Z, dD/dz and header OM of etc/dgrowth.fits (bin/make_boxes.py:307-313)."""
    etc = _etc_dir()
    if filename is None and etc and os.path.isfile(etc + "/dgrowth.fits"):
        filename = etc + "/dgrowth.fits"
    if filename is not None:
        d = fitsio.read(filename, ext=1)
        return d["Z"], d["dD/dz"], float(fitsio.read_header(filename, ext=1)["OM"])
    t = _npz()
    return t["dgrowth_Z"], t["dgrowth_dDdz"], float(t["dgrowth_OM"])


def planck_pk(filename=None):
    """K, PK and ZREF of etc/PlanckDR12.fits (py/SaclayMocks/powerspectrum.py:68-88)."""
    etc = _etc_dir()
    if filename is None and etc and os.path.isfile(etc + "/PlanckDR12.fits"):
        filename = etc + "/PlanckDR12.fits"
    if filename is not None:
        d = fitsio.read(filename, ext=1)
        return d["K"].copy(), d["PK"].copy(), float(fitsio.read_header(filename, ext=1)["ZREF"])
    t = _npz()
    return t["planck_K"].copy(), t["planck_PK"].copy(), float(t["planck_ZREF"])


def p1dmiss_tables():
    """(z[5], k[20000], pk[5,20000]) from etc/p1dmiss_z*.fits (cols k, P1DmissRSD)."""
    etc = _etc_dir()
    zs = [1.8, 2.2, 2.6, 3.0, 3.6]
    if etc and all(os.path.isfile(etc + "/p1dmiss_z%.1f.fits" % z) for z in zs):
        pk = []
        for z in zs:
            d = fitsio.read(etc + "/p1dmiss_z%.1f.fits" % z, ext=1)
            k = d["k"]
            pk.append(d["P1DmissRSD"])
        return np.array(zs), np.asarray(k), np.array(pk)
    t = _npz()
    return t["p1dmiss_z"], t["p1dmiss_k"], t["p1dmiss_pk"]


def nz_qso_desi():
    return _npz()["nz_qso_desi"]


def qso_lognormal_coef():
    """z, coef columns of etc/qso_lognormal_coef.txt (py/SaclayMocks/util.py:520-536)."""
    etc = _etc_dir()
    if etc and os.path.isfile(etc + "/qso_lognormal_coef.txt"):
        return np.loadtxt(etc + "/qso_lognormal_coef.txt")[:, :2]
    return _npz()["qso_lognormal_coef"]
