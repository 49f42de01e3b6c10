"""Quasar drawing — CPU restatement of bin/draw_qso.py:140-520 (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

One call = one x-slice (`-i`): the three lognormal boxes -> ptot(z) (draw_qso.py:228-251), the per-z-plane
selection cond1 & cond2 & cond3 with the legacy NumPy stream seeded `seed + i_slice` (:154, :394-467), the
redshift-space shift of the quasar redshift (:447-457) and the columns of QSO-<i>-<N>.fits (:493-521).
`-desi False` only (the footprint map etc/desi-healpix-weights.fits is a missing blob), `-random False`.
Pinned against the unmodified script run with stand-in wheels: tests/golden/ref_qso.npz.
"""
import numpy as np

from . import cosmology as co

N_QSO_EXP = 100.0         # constant.py:8
QSO_NZ_ADHOC = 0.213      # constant.py:9
RHO_SUM = 16452460        # constant.py:42


def qso_a_of_z(z, zb):
    """util.py:516-517."""
    return co.bias_qso(z) * (1 + zb) / (co.bias_qso(zb) * (1 + z))


def lognormal_coef_table():
    """util.py:520-536: (z, coef) with coef = 1 below the table (z = 0) and 0 above it (z = 10)."""
    d = co.tables()["qso_lognormal_coef"]
    return np.concatenate(([0.0], d[:, 0], [10.0])), np.concatenate(([1.0], d[:, 1], [0.0]))


def diffmod(a, b, c):
    """util.py:116-120."""
    d = (a - b) % c
    return np.minimum(d, c - d)


def compute_radec_r2(x, y, z, ra0, dec0):
    """box.py:201-237 (angles in radians)."""
    numra = np.cos(ra0) * x - np.sin(dec0) * np.sin(ra0) * y + np.cos(dec0) * np.sin(ra0) * z
    denomra = -np.sin(ra0) * x - np.sin(dec0) * np.cos(ra0) * y + np.cos(dec0) * np.cos(ra0) * z
    numdec = np.cos(dec0) * y + np.sin(dec0) * z
    R = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    ra = np.zeros(x.shape)
    with np.errstate(divide="ignore", invalid="ignore"):
        at = np.arctan(numra / denomra)
    m = (numra > 0) & (denomra > 0)
    ra[m] = at[m]
    m = (numra > 0) & (denomra < 0)
    ra[m] = at[m] + np.pi
    m = (numra < 0) & (denomra < 0)
    ra[m] = at[m] + np.pi
    m = (numra < 0) & (denomra > 0)
    ra[m] = at[m] + 2 * np.pi
    ra[(numra == 0) & (denomra > 0)] = 0
    ra[(numra > 0) & (denomra == 0)] = np.pi / 2
    ra[(numra == 0) & (denomra < 0)] = np.pi
    ra[(numra < 0) & (denomra == 0)] = 3 * np.pi / 2
    ra[(numra == 0) & (denomra == 0)] = 0
    dec = np.arcsin(numdec / R)
    return ra, dec, R


class Setup(object):
    """Everything of draw_qso.py that does not depend on the box values (:196-360): axes, limits, n(z) per cell,
    normalisation.  sigma_p = std of the three lognormal boxes of THIS slice (:191-194)."""

    def __init__(self, NXs, NY, NZ, NX_full, dcell, i_slice, nslice, ra0, dec0, dra, ddec, zmin, zmax, sigma_p,
                 dmax=3, cosmo=None):
        h = co.h
        self.NXs, self.NY, self.NZ, self.i_slice, self.nslice = NXs, NY, NZ, i_slice, nslice
        DX = DY = DZ = float(dcell)
        self.DX, self.DY, self.DZ = DX, DY, DZ
        self.ra0, self.dec0, self.dra, self.ddec = ra0, dec0, dra, ddec
        cosmo = cosmo or co.Cosmo()
        self.cosmo = cosmo
        LX = NXs * DX
        LX_full = NX_full * DX
        LY, LZ = NY * DY, NZ * DZ
        R0 = h * cosmo.r_comoving(co.z0)
        self.R0 = R0
        self.x_axis = (np.arange(NXs) + 0.5) * DX + (2 * i_slice - nslice) * LX_full / (2 * nslice)     # :208
        self.y_axis = (np.arange(NY) + 0.5) * DY - LY / 2
        self.z_axis = (np.arange(NZ) + 0.5) * DZ + R0 - LZ / 2
        self.z1, self.z2, self.z3 = co.z_QSO_bias
        self.coef_z, self.coef_v = lognormal_coef_table()
        Rmin, Rmax, tanx_max, tany_max = co.box_limit(LX_full, LY, LZ, R0, dmax * DX)                   # :254
        if zmin > 0:
            z_min = max(float(cosmo.r_2_z(Rmin / h)), zmin)
            Rmin = cosmo.r_comoving(z_min) * h
        else:
            z_min = float(cosmo.r_2_z(Rmin / h))
        if zmax > 0:
            z_max = min(float(cosmo.r_2_z(Rmax / h)), zmax)
            Rmax = cosmo.r_comoving(z_max) * h
        else:
            z_max = float(cosmo.r_2_z(Rmax / h))
        self.z_min, self.z_max = z_min, z_max
        if dra > 0 and ddec > 0:
            Ldec = np.sin(np.radians(dec0 + ddec)) - np.sin(np.radians(dec0 - ddec))
            surface = Ldec * 2 * np.radians(dra)
        else:
            surface = 4 * np.arctan(tanx_max) * np.arctan(tany_max)
        surfaceDeg = surface * (180 / np.pi) ** 2
        nQSOexp = N_QSO_EXP * surfaceDeg
        nQSOexp *= QSO_NZ_ADHOC
        nQSOexp /= nslice
        volFrac = (surface / 3) * (Rmax ** 3 - Rmin ** 3) / LX_full / LY / LZ
        # dN/dz per cell (:291-312)
        d = co.tables()["nz_qso_desi"]
        zlow, zhigh, dndz = d[:, 0], d[:, 1], d[:, 2]
        self.dz_interp = np.linspace((zlow[0] + zhigh[0]) / 2, (zlow[-1] + zhigh[-1]) / 2, 2 * NZ)
        dndz_interp = co.interp1d((zlow + zhigh) / 2, dndz, self.dz_interp)
        dn_cell = dndz_interp * DZ / cosmo.dist_hubble(self.dz_interp)
        dn_cell = dn_cell * DX * DY / cosmo.r_comoving(self.dz_interp) ** 2
        mmm = (self.dz_interp > z_min) & (self.dz_interp < z_max)
        self.sigma_p = tuple(float(s) for s in sigma_p)
        mean_rho = dn_cell[mmm] / self.cond1_correction(self.dz_interp[mmm])
        density_max, density_mean = np.max(mean_rho), np.mean(mean_rho)
        norm = nQSOexp / RHO_SUM
        norm *= density_max / density_mean
        norm /= volFrac
        if z_max > 2.1:
            N_zmin_zmax = dn_cell[(self.dz_interp > z_min) * (self.dz_interp < z_max)].sum()
            N_21_zmax = dn_cell[(self.dz_interp > 2.1) * (self.dz_interp < z_max)].sum()
            norm /= N_21_zmax / N_zmin_zmax
        self.norm, self.density_max = norm, density_max
        self.dn_cell = np.append(dn_cell, np.zeros(10 * NZ))                                            # :334
        dg_z, dg_v, _ = co.tables()["dgrowth_Z"], co.tables()["dgrowth_dDdz"], None
        self.dg_z, self.dg_v = dg_z, dg_v
        self.dgrowth0 = float(co.interp1d(dg_z, dg_v, 0.0))

    def coef(self, z):
        return co.interp1d(self.coef_z, self.coef_v, z)

    def cond1_correction(self, z):
        """The z dependence of <ptot> (draw_qso.py:315-325, :412-421)."""
        z1, z2, z3 = self.z1, self.z2, self.z3
        s1, s2, s3 = self.sigma_p
        c = self.coef(z)
        return (c * (np.exp((qso_a_of_z(z, z1) * s1) ** 2 / 2) * (z2 - z) / (z2 - z1)
                     + np.exp((qso_a_of_z(z, z2) * s2) ** 2 / 2) * (z - z1) / (z2 - z1))
                + (1 - c) * (np.exp((qso_a_of_z(z, z2) * s2) ** 2 / 2) * (z3 - z) / (z3 - z2)
                             + np.exp((qso_a_of_z(z, z3) * s3) ** 2 / 2) * (z - z2) / (z3 - z2)))


def ptot_box(st, b1, b2, b3):
    """draw_qso.py:197-199, 228-251: ptot = coef(z) (p12 - p23) + p23 from the three lognormal boxes (float32)."""
    h = co.h
    p1, p2, p3 = (np.exp(np.asarray(b, dtype=np.float32)) for b in (b1, b2, b3))        # float32 exp, in place there
    z_box = st.cosmo.r_2_z(np.sqrt((st.x_axis ** 2).reshape(-1, 1, 1) + (st.y_axis ** 2).reshape(-1, 1)
                                   + st.z_axis ** 2) / h)
    z1, z2, z3 = st.z1, st.z2, st.z3
    p1 = np.power(p1, qso_a_of_z(z_box, z1)).astype(np.float32)      # `p1 **= a` on a float32 array: float64 loop,
    p2 = np.power(p2, qso_a_of_z(z_box, z2)).astype(np.float32)      # result cast back to float32
    p3 = np.power(p3, qso_a_of_z(z_box, z3)).astype(np.float32)
    p12 = p1 * (z2 - z_box) / (z2 - z1) + p2 * (z_box - z1) / (z2 - z1)
    p23 = p2 * (z3 - z_box) / (z3 - z2) + p3 * (z_box - z2) / (z3 - z2)
    return st.coef(z_box) * (p12 - p23) + p23


def draw_uniforms(seed, NXs, NY, NZ):
    """The five legacy-NumPy draws of every z plane in the reference's order (draw_qso.py:425-445), as raw
    uniforms in [0, 1): u1, u2 [NZ, NXs, NY]; ux [NZ, NXs]; uy [NZ, NY]; uz [NZ, NXs, NY].  np.random.uniform(lo, hi)
    is lo + (hi - lo) * u with the same u.  Returns the RandomState as well (MJD / FIBERID are drawn after the loop)."""
    rs = np.random.RandomState(seed)
    u1 = np.empty((NZ, NXs, NY))
    u2 = np.empty((NZ, NXs, NY))
    ux = np.empty((NZ, NXs))
    uy = np.empty((NZ, NY))
    uz = np.empty((NZ, NXs, NY))
    for mz in range(NZ):
        u1[mz] = rs.random_sample((NXs, NY))
        u2[mz] = rs.random_sample((NXs, NY))
        ux[mz] = rs.random_sample(NXs)
        uy[mz] = rs.random_sample(NY)
        uz[mz] = rs.random_sample((NXs, NY))
    return (u1, u2, ux, uy, uz), rs


def draw_qso_slice(boxes, NX_full, dcell, i_slice, nslice, chunk, ra0, dec0, dra, ddec, zmin, zmax, seed, rsd=True,
                   dmax=3, uniforms=None):
    """boxes: dict boxln_1, boxln_2, boxln_3 (, vx, vy, vz) -> float32 [NXs, NY, NZ] of this slice.
    seed: the `-seed` argument (the stream is seeded seed + i_slice, draw_qso.py:154).
    Returns the columns of QSO-<i>-<N>.fits (dict of arrays) plus 'cells' = (mz, ix, iy) of every quasar and
    'nn_cond1' (the "QSOs in the full box" count)."""
    h, H0 = co.h, co.H0
    b1, b2, b3 = boxes["boxln_1"], boxes["boxln_2"], boxes["boxln_3"]
    NXs, NY, NZ = b1.shape
    sigma_p = (np.std(b1), np.std(b2), np.std(b3))
    st = Setup(NXs, NY, NZ, NX_full, dcell, i_slice, nslice, ra0, dec0, dra, ddec, zmin, zmax, sigma_p, dmax=dmax)
    ptot = ptot_box(st, b1, b2, b3)
    if uniforms is None:
        (u1, u2, ux, uy, uz), rs = draw_uniforms(seed + i_slice, NXs, NY, NZ)
    else:
        (u1, u2, ux, uy, uz), rs = uniforms
    cosmo = st.cosmo
    DX, DY, DZ = st.DX, st.DY, st.DZ
    out = {k: [] for k in ("z", "zrsd", "ra", "dec", "xx", "yy", "zz", "mz", "ix", "iy")}
    nn = 0
    delta_z = st.dz_interp[1] - st.dz_interp[0]
    for mz in range(NZ):
        XX, YY = st.x_axis, st.y_axis
        XY2 = (XX * XX).reshape(-1, 1) + YY * YY
        ZZ = st.z_axis[mz]
        RR = np.sqrt(ZZ * ZZ + XY2)
        redshift = cosmo.r_2_z(RR / h)
        iz = ((redshift - st.dz_interp[0]) / delta_z).round().astype(int)
        density = st.dn_cell[iz] / st.cond1_correction(redshift)
        cond1 = u1[mz] < st.norm * ptot[:, :, mz]
        nn += int(cond1.sum())
        cond2 = st.density_max * u2[mz] < density
        XX = XX + (-DX / 2 + (DX / 2 - -DX / 2) * ux[mz])
        YY = YY + (-DY / 2 + (DY / 2 - -DY / 2) * uy[mz])
        XXX = XX.reshape(-1, 1) * np.ones(NY)
        YYY = np.ones(NXs).reshape(-1, 1) * YY
        ZZZ = ZZ + (-DZ / 2 + (DZ / 2 - -DZ / 2) * uz[mz])
        ra, dec, RR = compute_radec_r2(XXX, YYY, ZZZ, np.radians(ra0), np.radians(dec0))
        ra, dec = np.degrees(ra), np.degrees(dec)
        redshift = cosmo.r_2_z(RR / h)
        if redshift.min() > st.z_max + 1.:
            continue
        if rsd:
            vpar = (XXX * boxes["vx"][:, :, mz] + YYY * boxes["vy"][:, :, mz] + ZZZ * boxes["vz"][:, :, mz]) / RR
            msk = redshift < st.z_max + 1.
            RR_RSD = RR.copy()
            RR_RSD[msk] += vpar[msk] * (1 + redshift[msk]) * co.interp1d(st.dg_z, st.dg_v, redshift[msk]) / (
                st.dgrowth0 * H0)
            redshift_RSD = cosmo.r_2_z(RR_RSD / h)
        else:
            redshift_RSD = redshift
        cond3 = ((diffmod(ra, ra0, 360) < dra) * (diffmod(dec, dec0, 180) < ddec) * (redshift_RSD > st.z_min)
                 * (redshift_RSD < st.z_max))
        iq = np.where(cond1 * cond2 * cond3)
        if iq[0].size == 0:
            continue
        out["xx"].append(XX[iq[0]])
        out["yy"].append(YY[iq[1]])
        out["zz"].append(ZZZ[iq])
        out["z"].append(redshift[iq])
        out["zrsd"].append(redshift_RSD[iq])
        out["ra"].append(ra[iq])
        out["dec"].append(dec[iq])
        out["mz"].append(np.full(iq[0].size, mz))
        out["ix"].append(iq[0])
        out["iy"].append(iq[1])
    cat = {k: (np.concatenate(v) if v else np.zeros(0)) for k, v in out.items()}
    n = len(cat["z"])
    thing_id = (chunk * 1e9 + i_slice * 1e6 + np.arange(n) + 1).astype(int)                  # :495
    mjd = rs.randint(51608, high=57521, size=n)
    fiberid = rs.randint(1, high=1001, size=n)
    pmf = np.array(["%d-%d-%d" % (t, m, f) for t, m, f in zip(thing_id, mjd, fiberid)], dtype="S21")
    return {"Z_QSO_NO_RSD": np.float32(cat["z"]), "Z_QSO_RSD": np.float32(cat["zrsd"]), "RA": np.float32(cat["ra"]),
            "DEC": np.float32(cat["dec"]), "HDU": np.int32(np.ones(n) * i_slice), "THING_ID": thing_id,
            "PLATE": thing_id, "MJD": np.int32(mjd), "FIBERID": np.int32(fiberid), "PMF": pmf,
            "XX": np.float32(cat["xx"]), "YY": np.float32(cat["yy"]), "ZZ": np.float32(cat["zz"]),
            "cells": np.stack([cat["mz"], cat["ix"], cat["iy"]], axis=1).astype(np.int64) if n else np.zeros((0, 3), int),
            "nn_cond1": nn, "setup": st, "ptot": ptot,
            "f64": {k: cat[k] for k in ("z", "zrsd", "ra", "dec", "xx", "yy", "zz")}}
