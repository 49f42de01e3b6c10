"""Skewer extraction — restated from bin/make_spectra.py:39-139 (ComputeWeight, computeRho,
ReadSpec), :196-295 (slab loader), :300-355 (pixel grid, QSO-file pre-selection),
:412-522 (per-QSO line of sight).

One documented deviation: the reference's gather has no bounds check (flat index
ny*nz*lx + nz*ly + lz, make_spectra.py:111); here every neighbour index is clamped to the
slab, which is identical whenever the 7^3 window is inside the slab (always true for
sightlines that respect the box_limit margin)."""
import numpy as np
from numba import njit

from . import cosmology as co

FIELDS = ("box", "eta_xx", "eta_yy", "eta_zz", "eta_xy", "eta_xz", "eta_yz", "vx", "vy", "vz")


@njit(cache=True)
def read_spec(rho, exx, eyy, ezz, exy, exz, eyz, vx, vy, vz, nx, ny, nz, Xvec, XvecSlice, Yvec, Zvec,
              LX, LY, LZ, DX, DY, DZ, R0, dmax, rsd, dla, imin, imax):
    """make_spectra.py:90-139 with ComputeWeight (:40-62) and computeRho (:66-68) inlined.
    Fields are the raveled slab arrays [nx, ny, nz]; float64 weights and accumulation."""
    n = XvecSlice.size
    spectrum = np.full(n, -1000000.0)
    eta_par = np.zeros(n)
    vpar = np.zeros(n)
    imax = min(imax, n)
    sig2 = 2 * DX * DX
    for p in range(imin, imax):
        X = XvecSlice[p]
        Xt = Xvec[p]
        Y = Yvec[p]
        Z = Zvec[p]
        ix = int((X + LX / 2) / DX)
        iy = int((Y + LY / 2) / DY)
        iz = int((Z + LZ / 2 - R0) / DZ)
        cx = (ix + 0.5) * DX - LX / 2
        cy = (iy + 0.5) * DY - LY / 2
        cz = (iz + 0.5) * DZ - LZ / 2 + R0
        sw = 0.0
        s0 = 0.0
        sxx = syy = szz = sxy = sxz = syz = 0.0
        svx = svy = svz = 0.0
        for a in range(-dmax, dmax + 1):
            dx2 = (a * DX + cx - X) ** 2
            la = min(max(ix + a, 0), nx - 1)
            for b in range(-dmax, dmax + 1):
                dy2 = (b * DY + cy - Y) ** 2
                lb = min(max(iy + b, 0), ny - 1)
                for c in range(-dmax, dmax + 1):
                    dz2 = (c * DZ + cz - Z) ** 2
                    lc = min(max(iz + c, 0), nz - 1)
                    w = np.exp(-(dx2 + dy2 + dz2) / sig2)
                    j = ny * nz * la + nz * lb + lc
                    sw += w
                    s0 += w * rho[j]
                    if rsd:
                        sxx += w * exx[j]
                        syy += w * eyy[j]
                        szz += w * ezz[j]
                        sxy += w * exy[j]
                        sxz += w * exz[j]
                        syz += w * eyz[j]
                        if dla:
                            svx += w * vx[j]
                            svy += w * vy[j]
                            svz += w * vz[j]
        spectrum[p] = s0 / sw
        if rsd:
            RR = Xt ** 2 + Y ** 2 + Z ** 2
            exx_ = sxx / sw
            eyy_ = syy / sw
            ezz_ = szz / sw
            exy_ = sxy / sw
            exz_ = sxz / sw
            eyz_ = syz / sw
            eta_par[p] = (Xt * exx_ * Xt + Y * eyy_ * Y + Z * ezz_ * Z + 2 * Xt * exy_ * Y
                          + 2 * Xt * exz_ * Z + 2 * Y * eyz_ * Z) / RR
            if dla:
                vpar[p] = (svx / sw * Xt + svy / sw * Y + svz / sw * Z) / np.sqrt(RR)
    return spectrum, eta_par, vpar


def slab_planes(islice, nslice, NX, dmax=3):
    """make_spectra.py:220-221."""
    ixmin = max((islice * NX) // nslice - dmax, 0)
    ixmax = min(((islice + 1) * NX) // nslice + dmax, NX)
    return ixmin, ixmax


class Geometry(object):
    """Box + cosmology set-up shared by all slices (make_spectra.py:196-206, 300-321)."""

    def __init__(self, NX, NY, NZ, dcell, zmin=1.8, zmax=3.6, pixel=0.2, dmax=3):
        self.NX, self.NY, self.NZ = NX, NY, NZ
        self.DX = self.DY = self.DZ = dcell
        self.LX, self.LY, self.LZ = dcell * NX, dcell * NY, dcell * NZ
        self.dmax, self.pixel, self.zmin, self.zmax = dmax, pixel, zmin, zmax
        self.cosmo = co.Cosmo()
        self.R0 = co.h * float(self.cosmo.r_comoving(co.z0))
        self.R_vec, self.lambda_vec = co.pixel_grid(self.cosmo, zmin, zmax, pixel)
        self.npixeltot = len(self.lambda_vec)
        t = co.tables()
        self.dg_z, self.dg = t["dgrowth_Z"], t["dgrowth_dDdz"]
        self.dgrowth0 = float(co.interp1d(self.dg_z, self.dg, 0.0))


def make_spectra_slice(geom, boxes, qso_files, islice, nslice, ra0, dec0, rsd=True, dla=True):
    """One make_spectra.py process (-i islice -N nslice).

    boxes: dict name -> float32 [NX,NY,NZ] full boxes (the FITS files of make_boxes re-assembled).
    qso_files: list (length nslice) of structured arrays with RA, DEC, Z_QSO_RSD, Z_QSO_NO_RSD, THING_ID, HDU.
    Returns a list of pieces: dict(id, hdu, ra, dec, z, z_norsd, lam, delta_l, eta_par, velo_par, redshift).
    """
    g = geom
    NX, dmax, DX = g.NX, g.dmax, g.DX
    ixmin, ixmax = slab_planes(islice, nslice, NX, dmax)
    nxs = ixmax - ixmin
    sl = {k: np.ascontiguousarray(boxes[k][ixmin:ixmax]).ravel() for k in FIELDS if k in boxes}
    empty = np.zeros(0, dtype=np.float32)
    f = [sl.get(k, empty) for k in FIELDS]
    xSlicemin = g.LX * islice / nslice - g.LX / 2          # make_spectra.py:279-280
    xSlicemax = g.LX * (islice + 1) / nslice - g.LX / 2
    # QSO-file half selection and conservative tan(x) cut (make_spectra.py:324-355)
    if islice >= nslice // 2:
        ifile0, ifile1 = nslice // 2, nslice
        tanx_slice_max = (DX * (NX / nslice) * (islice + 1) - g.LX / 2) / (g.R0 - g.LZ / 2)
    else:
        ifile0, ifile1 = 0, nslice // 2
        tanx_slice_max = np.abs((DX * (NX / nslice) * islice - g.LX / 2) / (g.R0 - g.LZ / 2))
    qsos = np.concatenate([qso_files[i] for i in range(ifile0, ifile1)])
    pieces = []
    cosmo = g.cosmo
    for q in qsos:
        ra, dec = q["RA"], q["DEC"]
        zQSO = q["Z_QSO_RSD"] if rsd else q["Z_QSO_NO_RSD"]
        R_QSO = co.h * cosmo.r_comoving(zQSO)
        X_QSO, Y_QSO, Z_QSO = co.compute_xyz2(np.radians(ra), np.radians(dec), R_QSO, np.radians(ra0), np.radians(dec0))
        if np.abs(X_QSO / Z_QSO) > tanx_slice_max:
            continue
        if (zQSO < g.zmin) | (zQSO > g.zmax):
            continue
        cut = (g.R_vec * X_QSO / R_QSO > xSlicemin)
        Rvec = g.R_vec[cut]
        mylambda = g.lambda_vec[cut]
        cut = (Rvec * X_QSO / R_QSO <= xSlicemax)
        Rvec = Rvec[cut]
        mylambda = mylambda[cut]
        if len(Rvec) < 1:
            continue
        redshift = cosmo.r_2_z(Rvec / co.h)
        Xvec = Rvec * X_QSO / R_QSO
        Yvec = Rvec * Y_QSO / R_QSO
        Zvec = Rvec * Z_QSO / R_QSO
        XvecSlice = Xvec - g.LX * islice / nslice
        if islice > 0:
            XvecSlice += DX * dmax
        w = np.where(mylambda > co.lylimit * (1 + zQSO))[0]
        imin = len(mylambda) if len(w) == 0 else w[0]
        w = np.where(mylambda < co.lya * (1 + zQSO))[0]
        imax = 0 if len(w) == 0 else w[-1] + 1
        delta_l, eta_par, velo_par = read_spec(f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8], f[9],
                                               nxs, g.NY, g.NZ, Xvec, XvecSlice, Yvec, Zvec,
                                               g.LX, g.LY, g.LZ, g.DX, g.DY, g.DZ, g.R0, dmax,
                                               int(rsd), int(dla), int(imin), int(imax))
        if rsd and dla:                                           # make_spectra.py:510
            velo_par = velo_par * ((1 + redshift) * cosmo.dist_hubble(redshift) / cosmo.dist_hubble(0)
                                   * co.interp1d(g.dg_z, g.dg, redshift) / g.dgrowth0)
        pieces.append(dict(id=int(q["THING_ID"]), hdu=int(q["HDU"]), ra=ra, dec=dec, z=q["Z_QSO_RSD"],
                           z_norsd=q["Z_QSO_NO_RSD"], lam=np.float32(mylambda), delta_l=np.float32(delta_l),
                           eta_par=np.float32(eta_par), velo_par=np.float32(velo_par),
                           redshift=np.float32(redshift), npix_forest=int(max(imax - imin, 0))))
    return pieces
