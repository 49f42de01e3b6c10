"""Spectral weights sqrt(P(|k|)/Vcell) — restated from bin/interpolate_pk.py:17-26, 51-77,
98-150 and py/SaclayMocks/powerspectrum.py:68-200 (P_0, P_ln, xi_from_pk, pk_from_xi,
LogNormalP, with the py2 `nr = nk/2` repaired to an int)."""
import numpy as np
from scipy.interpolate import InterpolatedUnivariateSpline

from . import cosmology as co


def _planck_at_z0(G_times_bias=1.0):
    """powerspectrum.py:71-88 / 104-118: P(k) of etc/PlanckDR12.fits rescaled to z=0, zero prepended."""
    t = co.tables()
    k = np.append(np.arange(1), t["planck_K"])
    P = t["planck_PK"] / co.fgrowth(float(t["planck_ZREF"]), co.omega_M_0) ** 2
    P = np.append(np.arange(1), P) * G_times_bias ** 2
    return k, P


def xi_from_pk(k, pk, nk=32 * 1024, direct=True):
    """powerspectrum.py:151-176."""
    if k[0] != 0:
        k = np.append([0], k)
        pk = np.append([0], pk)
    spl = InterpolatedUnivariateSpline(k, pk)
    kmax = np.max(k)
    nk = int(nk)
    kIn = np.linspace(0, kmax, nk)
    pkIn = spl(kIn)
    r = 2. * np.pi * np.arange(nk) / kmax
    pkk = kIn * pkIn
    r[0] = 1E-10
    cric = -np.imag(np.fft.fft(pkk) / nk) / r / 2. / np.pi ** 2 * kmax
    r[0] = 0
    spl2 = InterpolatedUnivariateSpline(k, pk * k * k)
    cric[0] = spl2.integral(0, kmax) / 2 / np.pi ** 2
    return r[0:nk // 2], cric[0:nk // 2]


def pk_from_xi(r, xi, nr=32 * 1024):
    """powerspectrum.py:186-189."""
    k, Pk = xi_from_pk(r, xi, nr, direct=False)
    return k, Pk * (2 * np.pi) ** 3


def lognormal_p(k, P, nk=1024 * 1024):
    """powerspectrum.py:194-200."""
    r, xi = xi_from_pk(k, P, nk=nk)
    cln = np.log(1 + xi)
    kln, Pln = pk_from_xi(r, cln, nr=nk // 2)
    return kln, np.maximum(Pln, 0)


_spline_cache = {}


def spline_P0():
    if "P0" not in _spline_cache:
        _spline_cache["P0"] = InterpolatedUnivariateSpline(*_planck_at_z0())
    return _spline_cache["P0"]


def spline_Pln(i):
    """i in 0,1,2: lognormal P at z_QSO_bias[i] with G*b_QSO (interpolate_pk.py:98-104)."""
    key = "Pln%d" % i
    if key not in _spline_cache:
        z = co.z_QSO_bias[i]
        gb = co.fgrowth(z, co.omega_M_0) * co.bias_qso(z)
        _spline_cache[key] = InterpolatedUnivariateSpline(*lognormal_p(*_planck_at_z0(gb)))
    return _spline_cache[key]


def k_norm(NX, NY, NZ, dcell, x0=0, x1=None):
    """interpolate_pk.py:63-77: float32 |k| on rows kx[x0:x1] (fftfreq order)."""
    k_ny = np.pi / dcell
    kx = np.fft.fftfreq(NX) * 2 * k_ny
    ky = np.fft.fftfreq(NY) * 2 * k_ny
    kz = np.fft.rfftfreq(NZ) * 2 * k_ny
    kx = kx[x0:x1]
    kz = np.float32(kz)
    ky = np.float32(ky.reshape(-1, 1))
    kx = np.float32(kx.reshape(-1, 1, 1))
    return np.sqrt(kx * kx + ky * ky + kz * kz)


def weight_from_spline(spl, k, dcell):
    """interpolate_pk.py:17-26: float32(sqrt(float32(max(P(k),0))/Vcell))."""
    Vcell = np.float32(dcell ** 3)
    myP = np.float32(np.maximum(spl(k), 0))
    return np.float32(np.sqrt(myP / Vcell))


def weights(NX, NY, NZ, dcell, x0=0, x1=None):
    """dict Pln1,Pln2,Pln3,P0 -> float32 [x1-x0, NY, NZ/2+1] (the four HDUs of P<NX>-<NY>-<NZ>.fits)."""
    k = k_norm(NX, NY, NZ, dcell, x0, x1)
    out = {}
    for i in range(3):
        out["Pln%d" % (i + 1)] = weight_from_spline(spline_Pln(i), k, dcell)
    out["P0"] = weight_from_spline(spline_P0(), k, dcell)
    return out
