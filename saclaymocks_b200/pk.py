"""Spectral weights sqrt(P(|k|)/Vcell) of the four make_boxes products -- host side of the interpolate_pk /
merge_pk stage (bin/interpolate_pk.py:17-26, 63-77, 98-150; py/SaclayMocks/powerspectrum.py:68-200):
P0 is the linear Planck P(k) rescaled to z = 0; Pln1..3 are its lognormal transforms at z = 1.9, 2.75, 3.6 with
G(z) * b_QSO(z)."""
import numpy as np
from scipy.interpolate import InterpolatedUnivariateSpline

from . import constant
from . import tables
from .cosmo import fgrowth

_Z_LN = (constant.z_QSO_bias_1, constant.z_QSO_bias_2, constant.z_QSO_bias_3)
_splines = {}


def bias_qso(z):
    """py/SaclayMocks/util.py:508-513 (Laurent et al. 2017)."""
    return 3.7 * ((1 + z) / (1 + 2.33)) ** 1.7


def _input_pk(G_times_bias=1.0):
    K, PK, zref = tables.planck_pk()
    P = PK / fgrowth(zref, constant.omega_M_0) ** 2
    k = np.concatenate(([0.0], K))
    P = np.concatenate(([0.0], P)) * G_times_bias ** 2
    return k, P


def _hankel(k, pk, n, forward_scale, fft=None):
    """One direction of the P(k) <-> xi(r) pair on a uniform grid of n points up to max(k)
    (powerspectrum.py:151-176): xi(r) = -Im FFT[k P(k)] / (2 pi^2 r) * dk-normalisation."""
    spl = InterpolatedUnivariateSpline(k, pk)
    kmax = np.max(k)
    kin = np.linspace(0, kmax, n)
    r = 2. * np.pi * np.arange(n) / kmax
    integrand = kin * spl(kin)
    r[0] = 1e-10
    xi = -np.imag((fft or np.fft.fft)(integrand) / n) / r / 2. / np.pi ** 2 * kmax
    r[0] = 0
    xi[0] = InterpolatedUnivariateSpline(k, pk * k * k).integral(0, kmax) / 2 / np.pi ** 2
    return r[:n // 2], xi[:n // 2] * forward_scale


def lognormal_pk(k, P, nk=1024 * 1024, fft=None):
    """P(k) -> xi(r) -> ln(1 + xi) -> P_ln(k)   (powerspectrum.py:194-200).  fft: the transform to use (default
    np.fft.fft; gpu_fft() runs the two 2^20 / 2^19-point float64 transforms on the GPU, smk_fft1d_f64)."""
    r, xi = _hankel(k, P, nk, 1.0, fft)
    kln, Pln = _hankel(r, np.log(1 + xi), nk // 2, (2 * np.pi) ** 3, fft)
    return kln, np.maximum(Pln, 0)


def gpu_fft(device=None):
    """np.fft.fft replacement for 1-D float64 / complex128 arrays of power-of-two length on the GPU (smk_fft1d_f64)."""
    import ctypes as C
    import torch
    from . import _lib
    dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    ctx = _lib.StreamCtx(dev)

    def fft(a):
        x = torch.as_tensor(np.ascontiguousarray(a, dtype=np.complex128), device=dev)
        n = x.numel()
        out, work = torch.empty_like(x), torch.empty_like(x)
        _lib.check(ctx.lib.smk_fft1d_f64(ctx.handle(), n, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()),
                                         C.c_void_p(work.data_ptr())))
        return out.cpu().numpy()
    return fft


def spline(name):
    """Cubic interpolating spline of P0 / Pln1 / Pln2 / Pln3 (FITPACK, as the reference)."""
    if name not in _splines:
        if name == "P0":
            _splines[name] = InterpolatedUnivariateSpline(*_input_pk())
        else:
            z = _Z_LN[int(name[-1]) - 1]
            gb = fgrowth(z, constant.omega_M_0) * bias_qso(z)
            _splines[name] = InterpolatedUnivariateSpline(*lognormal_pk(*_input_pk(gb)))
    return _splines[name]


def weight(name, k, dcell):
    """float32(sqrt(float32(max(P(k), 0)) / Vcell)) for float32 |k| (interpolate_pk.py:17-26)."""
    Vcell = np.float32(dcell ** 3)
    p = np.float32(np.maximum(spline(name)(k), 0))
    return np.float32(np.sqrt(p / Vcell))


def k_norm(NX, NY, NZ, dcell, x0=0, x1=None):
    """float32 |k| for kx rows [x0, x1) in fftfreq order (interpolate_pk.py:63-77)."""
    k_ny = np.pi / dcell
    kx = np.float32((np.fft.fftfreq(NX) * 2 * k_ny)[x0:x1].reshape(-1, 1, 1))
    ky = np.float32((np.fft.fftfreq(NY) * 2 * k_ny).reshape(-1, 1))
    kz = np.float32(np.fft.rfftfreq(NZ) * 2 * k_ny)
    return np.sqrt(kx * kx + ky * ky + kz * kz)


def weight_tables(NX, NY, NZ, dcell, x0=0, x1=None):
    """The four HDUs of P<NX>-<NY>-<NZ>.fits for kx rows [x0, x1)."""
    k = k_norm(NX, NY, NZ, dcell, x0, x1)
    return {n: weight(n, k, dcell) for n in ("Pln1", "Pln2", "Pln3", "P0")}


def weight_curve(name, dcell, kmax, n=1 << 20):
    """sqrt(P(k)/Vcell) sampled on a uniform k grid (for fast table look-ups when exactness is not needed)."""
    k = np.linspace(0.0, kmax, n)
    return k, np.sqrt(np.maximum(spline(name)(k), 0) / dcell ** 3)


def ppoly(name):
    """Piecewise-polynomial form of spline(name): (breaks[n+1], coefs[4][n]) float64, for smk_pk_weights."""
    from scipy.interpolate import PPoly
    pp = PPoly.from_spline(spline(name)._eval_args)
    x, c = pp.x, pp.c
    keep = np.diff(x) > 0                       # drop the zero-length intervals of the repeated end knots
    return np.ascontiguousarray(np.concatenate((x[:-1][keep], x[-1:]))), np.ascontiguousarray(c[:, keep])
