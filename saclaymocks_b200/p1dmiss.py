"""P1D_missing tables for the small-scale 1-D field (bin/merge_spectra.py:105-118, 303-324).

The reference's default `etc/pkmiss_interp.fits.gz` is one of its missing large blobs
(/root/reference/.MISSING_LARGE_BLOBS).  `build_pkmiss_interp` writes a file of the same
layout (IMAGE HDUs `z`, `k`, `pk`, `sigma`; py/SaclayMocks/util.py:448-457) from the five
shipped `etc/p1dmiss_z*.fits`, with `sigma` from the util.sigma_p1d formula
(py/SaclayMocks/util.py:372-398)."""
import numpy as np

from . import fitsio_lite as fitsio
from . import tables


def sigma_p1d(k, pk_row, pixel=0.2, N=10000):
    """sigma of delta_s for one tabulated P1D_miss(k) row — py/SaclayMocks/util.py:372-398."""
    L = N * pixel
    kj = 2 * np.pi / L * np.arange(1, N / 2)
    var_s = 2 * np.interp(kj, k, pk_row).sum() / L
    var_s += np.interp(0.0, k, pk_row) / L
    var_s += np.interp(np.pi / pixel, k, pk_row) / L
    return np.sqrt(var_s)


def pkmiss_arrays(pixel=0.2):
    z, k, pk = tables.p1dmiss_tables()
    sigma = np.array([sigma_p1d(k, pk[i], pixel) for i in range(len(z))])
    return z, k, pk, sigma


def build_pkmiss_interp(filename, pixel=0.2):
    z, k, pk, sigma = pkmiss_arrays(pixel)
    f = fitsio.FITS(filename, "rw", clobber=True)
    f.write(z, extname="z")
    f.write(k, extname="k")
    f.write(pk, extname="pk")
    f.write(sigma, extname="sigma")
    f.close()
    return filename


class InterpP1Dmissing(object):
    """Nearest tabulated z, then linear in k — py/SaclayMocks/util.py:448-474."""

    def __init__(self, infile=None):
        if infile is None:
            self.z, self.k, self.pk, self.sigma = pkmiss_arrays()
        else:
            f = fitsio.FITS(infile)
            self.z, self.k, self.pk = f["z"].read(), f["k"].read(), f["pk"].read()
            self.sigma = f["sigma"].read()
        self.zmin, self.zmax = self.z.min(), self.z.max()

    def iz(self, redshift):
        return int(np.argsort(np.abs(self.z - redshift))[0])

    def __call__(self, redshift, k):
        return np.interp(k, self.k, self.pk[self.iz(redshift)])
