"""Cosmology tables and box geometry used to set up the skewer kernels (host side, tiny).
Mirrors py/SaclayMocks/util.py:123-134 (fgrowth), :300-314 (InterpFitsTable), :793-833 (cosmo)
and py/SaclayMocks/box.py:136-161, 240-250."""
import numpy as np

from . import constant


def lin_interp(x, y, xnew, what="value"):
    """Linear interpolation that refuses to extrapolate (scipy interp1d default, util.py:311-313)."""
    xn = np.asarray(xnew, dtype=np.float64)
    if xn.size and (xn.min() < x[0] or xn.max() > x[-1]):
        raise ValueError("{} out of the tabulated range [{}, {}]".format(what, x[0], x[-1]))
    i = np.clip(np.searchsorted(x, xn), 1, len(x) - 1)
    slope = (y[i] - y[i - 1]) / (x[i] - x[i - 1])
    return slope * (xn - x[i - 1]) + y[i - 1]


def fgrowth(z, Om0=constant.omega_M_0, unnormed=False):
    Om = 1 / (1 + (1 - Om0) / (Om0 * (1 + z) ** 3))
    Ol = 1 - Om
    a = 1 / (1 + z)
    norm = 1.0 if unnormed else 1.0 / fgrowth(0.0, Om0, unnormed=True)
    return norm * (5. / 2.) * a * Om / (Om ** (4. / 7.) - Ol + (1. + Om / 2.) * (1. + Ol / 70.))


class cosmo(object):
    """Flat-LCDM distance tables: 10 000 trapezoid steps to z = 10 (distances in Mpc, H0 in km/s/Mpc)."""

    def __init__(self, Om=constant.omega_M_0, Ok=constant.omega_k_0, H0=100 * constant.h):
        Ol = 1. - Ok - Om
        nbins, zmax = 10000, 10.
        dz = zmax / nbins
        self.z = np.arange(nbins) * dz
        self.hub = H0 * np.sqrt(Ol + Ok * (1. + self.z) ** 2 + Om * (1. + self.z) ** 3)
        step = constant.c * (1. / self.hub[:-1] + 1. / self.hub[1:]) / 2. * dz
        self.chi = np.concatenate(([0.], np.cumsum(step)))

    def r_comoving(self, z):
        return lin_interp(self.z, self.chi, z, "redshift")

    def r_2_z(self, r):
        return lin_interp(self.chi, self.z, r, "distance")

    def dist_hubble(self, z):
        return lin_interp(self.z, constant.c / self.hub, z, "redshift")


def ComputeXYZ2(ra, dec, R, ra0, dec0):
    """(ra, dec, R) -> box frame, angles in radians (box.py:240-250)."""
    # cos/sin run in the dtype of ra/dec (float32 for catalogue columns, as in the reference) and are then
    # promoted to float64 by the float64 factors
    f64 = np.float64
    cd, sd, cr, sr = f64(np.cos(dec)), f64(np.sin(dec)), f64(np.cos(ra)), f64(np.sin(ra))
    c0, s0, cr0, sr0 = np.cos(dec0), np.sin(dec0), np.cos(ra0), np.sin(ra0)
    x = R * (cr0 * cd * sr - sr0 * cd * cr)
    y = R * (-sr0 * s0 * cd * sr + c0 * sd - cr0 * s0 * cd * cr)
    z = R * (c0 * sr0 * cd * sr + s0 * sd + cr0 * c0 * cd * cr)
    return x, y, z


def box_limit(LX, LY, LZ, R0, margin):
    """box.py:136-161."""
    Rmax = R0 + LZ / 2 - margin
    sx = (LX / 2 - margin) / Rmax
    sy = (LY / 2 - margin) / Rmax
    smax = (np.sqrt(LX * LX + LY * LY) / 2 - margin) / Rmax
    Rmin = (R0 - LZ / 2 + margin) / np.sqrt(1 - smax ** 2)
    return Rmin, Rmax, sx / np.sqrt(1 - sx ** 2), sy / np.sqrt(1 - sy ** 2)
