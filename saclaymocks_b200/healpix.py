"""HEALPix ang2pix (NESTED and RING) — replaces `healpy.ang2pix` for util.radec2pix
(py/SaclayMocks/util.py:105-109), used by merge_spectra to name its output files
(bin/merge_spectra.py:220, 385).  Standard HEALPix projection (Gorski et al. 2005)."""
import numpy as np


def _spread_bits(v):
    v = v.astype(np.int64)
    out = np.zeros_like(v)
    for b in range(30):
        out |= ((v >> b) & 1) << (2 * b)
    return out


def _face_xy(nside, theta, phi):
    theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    phi = np.atleast_1d(np.asarray(phi, dtype=np.float64))
    z = np.cos(theta)
    za = np.abs(z)
    tt = np.mod(phi, 2 * np.pi) / (np.pi / 2)          # in [0,4)
    tt = np.where(tt >= 4.0, 0.0, tt)
    face = np.empty(z.shape, dtype=np.int64)
    ix = np.empty(z.shape, dtype=np.int64)
    iy = np.empty(z.shape, dtype=np.int64)
    eq = za <= 2.0 / 3.0
    # equatorial region
    t1 = nside * (0.5 + tt)
    t2 = nside * z * 0.75
    jp = np.floor(t1 - t2).astype(np.int64)            # ascending edge line index
    jm = np.floor(t1 + t2).astype(np.int64)            # descending edge line index
    ifp = jp // nside
    ifm = jm // nside
    f_eq = np.where(ifp == ifm, (ifp & 3) | 4, np.where(ifp < ifm, ifp & 3, (ifm & 3) + 8))
    ix_eq = jm & (nside - 1)
    iy_eq = nside - (jp & (nside - 1)) - 1
    # polar caps
    ntt = np.minimum(tt.astype(np.int64), 3)
    tp = tt - ntt
    tmp = nside * np.sqrt(3.0 * (1.0 - za))
    jpp = np.minimum((tp * tmp).astype(np.int64), nside - 1)
    jmp = np.minimum(((1.0 - tp) * tmp).astype(np.int64), nside - 1)
    north = z >= 0
    f_po = np.where(north, ntt, ntt + 8)
    ix_po = np.where(north, nside - jmp - 1, jpp)
    iy_po = np.where(north, nside - jpp - 1, jmp)
    face[:] = np.where(eq, f_eq, f_po)
    ix[:] = np.where(eq, ix_eq, ix_po)
    iy[:] = np.where(eq, iy_eq, iy_po)
    return face, ix, iy


_JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4])
_JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7])


def ang2pix(nside, theta, phi, nest=False):
    scalar = np.isscalar(theta) and np.isscalar(phi)
    face, ix, iy = _face_xy(nside, theta, phi)
    if nest:
        pix = face * nside * nside + _spread_bits(ix) + 2 * _spread_bits(iy)
    else:
        nl4 = 4 * nside
        jr = _JRLL[face] * nside - ix - iy - 1
        nr = np.where(jr < nside, jr, np.where(jr > 3 * nside, nl4 - jr, nside))
        n_before = np.where(jr < nside, 2 * nr * (nr - 1),
                            np.where(jr > 3 * nside, 12 * nside * nside - 2 * (nr + 1) * nr,
                                     2 * nside * (nside - 1) + (jr - nside) * nl4))
        kshift = np.where((jr < nside) | (jr > 3 * nside), 0, (jr - nside) & 1)
        jp = (_JPLL[face] * nr + ix - iy + 1 + kshift) // 2
        jp = np.where(jp > nl4, jp - nl4, np.where(jp < 1, jp + nl4, jp))
        pix = n_before + jp - 1
    return int(pix[0]) if scalar else pix


def radec2pix(nside, ra, dec, nest=True):
    """ra, dec in degrees (py/SaclayMocks/util.py:105-109)."""
    phi = np.asarray(ra) * np.pi / 180
    theta = np.pi / 2 - np.asarray(dec) * np.pi / 180
    return ang2pix(nside, theta, phi, nest=nest)
