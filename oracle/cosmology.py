"""Constants, cosmology tables and geometry — restated from
py/SaclayMocks/constant.py (whole file), py/SaclayMocks/util.py:123-134 (fgrowth),
:300-314 (InterpFitsTable), :793-833 (cosmo), py/SaclayMocks/box.py:136-161 (box_limit),
:240-250 (ComputeXYZ2)."""
import os

import numpy as np

c_kms = 299792.458            # scipy.constants.speed_of_light / 1000   (constant.py:6)
lya = 1215.67                 # constant.py:12
lylimit = 0.0                 # constant.py:14
lambda_min = 3476.0           # constant.py:18
h = 0.6731                    # constant.py:22
omega_M_0 = 0.31457
omega_lambda_0 = 0.68543
omega_k_0 = 0.0
z_QSO_bias = (1.9, 2.75, 3.6)  # constant.py:31-33
z0 = 1.70975268202            # constant.py:34
H0 = 100.0                    # constant.py:37

_NPZ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "saclaymocks_b200", "data", "etc_tables.npz")
_tab = None


def tables():
    global _tab
    if _tab is None:
        _tab = dict(np.load(_NPZ))
    return _tab


def interp1d(x, y, xnew):
    """scipy.interpolate.interp1d(kind='linear') arithmetic: slope*(x_new-x_lo)+y_lo."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    xn = np.asarray(xnew, dtype=np.float64)
    if xn.size and (xn.min() < x[0] or xn.max() > x[-1]):
        raise ValueError("interp1d: value out of range")
    idx = np.clip(np.searchsorted(x, xn), 1, len(x) - 1)
    lo, hi = idx - 1, idx
    slope = (y[hi] - y[lo]) / (x[hi] - x[lo])
    return slope * (xn - x[lo]) + y[lo]


def fgrowth(z, Om0=omega_M_0, unnormed=False):
    """util.py:123-134 (flat LCDM growth factor, Carroll-Press-Turner form)."""
    Om = 1 / (1 + (1 - Om0) / (Om0 * (1 + z) ** 3))
    Ol = 1 - Om
    a = 1 / (1 + z)
    norm = 1.0 if unnormed else 1.0 / fgrowth(0.0, Om0, unnormed=True)
    return norm * (5. / 2.) * a * Om / (Om ** (4. / 7.) - Ol + (1. + Om / 2.) * (1. + Ol / 70.))


def bias_qso(z):
    """util.py:508-513."""
    return 3.7 * ((1 + z) / (1 + 2.33)) ** 1.7


class Cosmo(object):
    """util.py:793-833: 10 000-step trapezoid chi(z) to z=10, linear interpolation. Distances in Mpc."""

    def __init__(self, Om=omega_M_0, Ok=omega_k_0, H0=100 * h):
        Ol = 1. - Ok - Om
        nbins, zmax = 10000, 10.
        dz = zmax / nbins
        z = np.arange(nbins) * dz
        hubble = H0 * np.sqrt(Ol + Ok * (1. + z) ** 2 + Om * (1. + z) ** 3 + 0. * (1. + z) ** 4)
        incr = c_kms * (1. / hubble[:-1] + 1. / hubble[1:]) / 2. * dz
        chi = np.concatenate(([0.], np.cumsum(incr)))          # same left-to-right sum as the reference loop
        self.z, self.chi, self.hubble = z, chi, hubble

    def r_comoving(self, z):
        return interp1d(self.z, self.chi, z)

    def r_2_z(self, r):
        return interp1d(self.chi, self.z, r)

    def dist_hubble(self, z):
        return interp1d(self.z, c_kms / self.hubble, z)


def compute_xyz2(ra, dec, R, ra0, dec0):
    """box.py:240-250 — angles in radians."""
    x = R * (np.cos(ra0) * np.cos(dec) * np.sin(ra) - np.sin(ra0) * np.cos(dec) * np.cos(ra))
    y = R * (-np.sin(ra0) * np.sin(dec0) * np.cos(dec) * np.sin(ra) + np.cos(dec0) * np.sin(dec)
             - np.cos(ra0) * np.sin(dec0) * np.cos(dec) * np.cos(ra))
    z = R * (np.cos(dec0) * np.sin(ra0) * np.cos(dec) * np.sin(ra) + np.sin(dec0) * np.sin(dec)
             + np.cos(ra0) * np.cos(dec0) * np.cos(dec) * np.cos(ra))
    return x, y, z


def box_limit(LX, LY, LZ, R0, margin):
    """box.py:136-161."""
    Rmax = R0 + LZ / 2 - margin
    sinx_max = (LX / 2 - margin) / Rmax
    tanx_max = sinx_max / np.sqrt(1 - sinx_max ** 2)
    siny_max = (LY / 2 - margin) / Rmax
    tany_max = siny_max / np.sqrt(1 - siny_max ** 2)
    sin_max = (np.sqrt(LX * LX + LY * LY) / 2 - margin) / Rmax
    Rmin = (R0 - LZ / 2 + margin) / np.sqrt(1 - sin_max ** 2)
    return Rmin, Rmax, tanx_max, tany_max


def pixel_grid(cosmo, zmin=1.8, zmax=3.6, pixel=0.2):
    """bin/make_spectra.py:300-321 — global comoving pixel grid and wavelengths after the lambda_min cut."""
    Rmin = h * cosmo.r_comoving(zmin)
    Rmax = h * cosmo.r_comoving(zmax)
    npixeltot = int((Rmax - Rmin) / pixel + 0.5)
    R_vec = Rmin + np.arange(npixeltot) * pixel
    lambda_vec = lya * (1 + cosmo.r_2_z(R_vec / h))
    cut = lambda_vec > lambda_min
    return R_vec[cut], lambda_vec[cut]
