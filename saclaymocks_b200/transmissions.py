"""DESI transmission files: transmission-<nside>-<pix>.fits.gz with HDUs METADATA,
WAVELENGTH, TRANSMISSION (and DLA) -- the on-disk format DESI tools consume, bin/make_transmissions.py:74-164 of the
reference.  `write_pixel` writes one file from arrays (host copies of the device spectra rows or the contents of
spectra_merged files); `from_merged_files` is the reference's regrouping of spectra_merged-<pix>-<hdu>.fits.gz."""
import glob
import os

import numpy as np

from . import constant
from . import fitsio_lite as fitsio

NAMES = ["RA", "DEC", "Z_noRSD", "Z", "MOCKID"]
DLA_NAMES = ["Z_DLA_NO_RSD", "Z_DLA_RSD", "N_HI_DLA", "MOCKID", "DLAID"]


def header_list(nside, nest=True):
    """make_transmissions.py:74-79."""
    return [{"name": "HPXNSIDE", "value": nside, "comment": "healpix nside parameter"},
            {"name": "HPXNEST", "value": "T" if nest else "F", "comment": "healpix scheme"},
            {"name": "OL", "value": constant.omega_lambda_0, "comment": "Omega_lambda_0"},
            {"name": "OM", "value": constant.omega_M_0, "comment": "Omega_M_0"},
            {"name": "OK", "value": constant.omega_k_0, "comment": "Omega_k_0"},
            {"name": "H0", "value": constant.h * 100, "comment": "H0"}]


def file_name(outDir, nside, pix):
    """make_transmissions.py:85: <outDir>/<pix//100>/<pix>/transmission-<nside>-<pix>.fits.gz."""
    return outDir + "/{}/{}/transmission-{}-{}.fits.gz".format(pix // 100, pix, nside, pix)


def write_pixel(fname, pix, nside, ra, dec, z_norsd, z, mockid, wavelength, flux, nest=True, dla_table=None):
    """One transmission file (make_transmissions.py:153-160)."""
    os.makedirs(os.path.dirname(fname), exist_ok=True)
    out = fitsio.FITS(fname, "rw", clobber=True)
    table = [np.float32(ra), np.float32(dec), np.float32(z_norsd), np.float32(z), np.asarray(mockid)]
    out.write(table, names=NAMES, header=header_list(nside, nest), extname="METADATA")
    out[-1].write_key("HPXPIXEL", int(pix), comment="Healpix pixel")
    out[-1].write_key("NSIDE", nside, comment="Healpix parameter")
    out.write(np.float32(wavelength), extname="WAVELENGTH")
    out.write(np.float32(flux), extname="TRANSMISSION")
    if dla_table is not None:
        out.write(list(dla_table), names=DLA_NAMES, extname="DLA")
    out.close()


def dla_rows(dla_cat, pix, ids):
    """DLAs of the quasars of one healpix pixel, in quasar order (make_transmissions.py:131-150)."""
    cat = dla_cat[dla_cat["PIXNUM"] == pix]
    cols = {k: [] for k in DLA_NAMES}
    for i in ids:
        msk = np.where(cat["MOCKID"] == i)[0]
        if len(msk) < 1:
            continue
        for k in DLA_NAMES:
            cols[k].append(cat[k][msk])
    if cols["MOCKID"]:
        c = {k: np.concatenate(v) for k, v in cols.items()}
        return [np.float32(c["Z_DLA_NO_RSD"]), np.float32(c["Z_DLA_RSD"]), np.float32(c["N_HI_DLA"]), c["MOCKID"],
                c["DLAID"]]
    x = np.array([], dtype=np.float32)
    return [x, x, x, x, x]


def pixel_of_file(f):
    """make_transmissions.py:64-67: the healpix index between 'spectra_merged-' and the next '-'."""
    i = f.find("spectra_merged-") + 15
    j = i + f[i:].find("-")
    return int(f[i:j])


def from_merged_files(inDir, outDir, pixels, nside=16, nest=True, dla_cat=None):
    """Regroup <inDir>/*/spectra_merged/spectra_merged-<pix>-<hdu>.fits.gz by healpix pixel
    (make_transmissions.py:57-160).  Files of a pixel are read in sorted order (the reference takes glob order).
    Returns the number of spectra_merged files consumed."""
    found = sorted(set(pixel_of_file(f) for f in glob.glob(inDir + "/*/spectra_merged/spectra_merged*")) & set(pixels))
    cpt = 0
    for pix in found:
        ra, dec, zn, zr, ids, fluxes, wavelength = [], [], [], [], [], [], None
        for f in sorted(glob.glob(inDir + "/*/spectra_merged/spectra_merged-{}-*".format(pix))):
            try:
                fits = fitsio.FITS(f)
                data = fits["METADATA"].read()
            except Exception:
                print("*WARNING* Fits file {} cannot be read".format(f))
                continue
            ra.append(data["RA"]), dec.append(data["DEC"]), zn.append(data["Z_noRSD"]), zr.append(data["Z"])
            ids.append(data["THING_ID"])
            if wavelength is None:
                wavelength = fits["LAMBDA"].read()
            fluxes.append(fits["FLUX"].read())
            fits.close()
        if not ra:
            continue
        cpt += len(ra)
        ids = np.concatenate(ids)
        write_pixel(file_name(outDir, nside, pix), pix, nside, np.concatenate(ra), np.concatenate(dec),
                    np.concatenate(zn), np.concatenate(zr), ids, wavelength, np.concatenate(fluxes), nest=nest,
                    dla_table=dla_rows(dla_cat, pix, ids) if dla_cat is not None else None)
    return cpt


def from_rows(outDir, ra, dec, z_norsd, z, mockid, wavelength, flux, nside=16, nest=True):
    """Transmission files straight from spectra rows (e.g. ChunkPipeline.out copied to the host): group the quasars by
    healpix pixel (nside 16 nested, merge_spectra.py:220) and write one file per pixel, quasars in MOCKID order."""
    from .healpix import radec2pix
    pix = radec2pix(nside, np.asarray(ra, dtype=np.float64), np.asarray(dec, dtype=np.float64), nest=nest)
    files = []
    for p in np.unique(pix):
        rows = np.where(pix == p)[0]
        rows = rows[np.argsort(np.asarray(mockid)[rows], kind="stable")]
        fname = file_name(outDir, nside, int(p))
        write_pixel(fname, int(p), nside, np.asarray(ra)[rows], np.asarray(dec)[rows], np.asarray(z_norsd)[rows],
                    np.asarray(z)[rows], np.asarray(mockid)[rows], wavelength, np.asarray(flux)[rows], nest=nest)
        files.append(fname)
    return files
