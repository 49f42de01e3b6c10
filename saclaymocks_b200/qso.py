"""Quasar drawing on device-resident boxes: host side of smk_draw_qso (mirrors bin/draw_qso.py of the reference).

`QsoSetup` is the part of draw_qso.py:196-360 that does not depend on the box values (axes, redshift limits, n(z) per
cell, the cond1 normalisation); `QsoDrawer.draw` runs the cell loop (draw_qso.py:228-251, 394-480) on the GPU for one
x-slab and returns the columns of QSO-<i>-<N>.fits (draw_qso.py:493-521).  With `uniforms=legacy_uniforms(seed + i)`
the selection reproduces the reference's quasars exactly (same cells, same sub-cell positions); with uniforms=None the
kernel draws Philox4x32-10 keyed by the global cell index.  `-desi False` / `-random False` only: the DESI footprint
map (etc/desi-healpix-weights.fits) is not distributed with the reference.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import constant
from . import cosmo as cosmo_mod
from . import tables

N_QSO_EXP = 100.0         # constant.py:8
QSO_NZ_ADHOC = 0.213      # constant.py:9
RHO_SUM = 16452460        # constant.py:42: sum of ptot over one slice of the nominal box (2560 x 2560 x 1536 / 512)
NOMINAL_SLICE_CELLS = 2560 * 2560 * 1536 // 512


def bias_qso(z):
    """util.py:508-513."""
    return 3.7 * ((1 + z) / (1 + 2.33)) ** 1.7


def qso_a_of_z(z, zb):
    """util.py:516-517."""
    return bias_qso(z) * (1 + zb) / (bias_qso(zb) * (1 + z))


def lognormal_coef():
    """util.py:520-536: z and coefficient of etc/qso_lognormal_coef.txt, 1 below the table and 0 above."""
    d = tables.qso_lognormal_coef()
    return np.concatenate(([0.0], d[:, 0], [10.0])), np.concatenate(([1.0], d[:, 1], [0.0]))


def _interp(x, y, v):
    return cosmo_mod.lin_interp(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64), v)


class QsoParams(C.Structure):
    """struct smk_qso_params of include/smk.h."""
    _d, _p, _i = C.c_double, C.c_void_p, C.c_int
    _fields_ = [("nxs", _i), ("ny", _i), ("nz", _i), ("ix0", _i), ("nx_full", _i),
                ("boxln", _p * 3), ("velo", _p * 3), ("rsd", _i),
                ("x_axis", _p), ("y_axis", _p), ("z_axis", _p), ("dx", _d), ("dy", _d), ("dz", _d),
                ("chi", _p), ("zt", _p), ("ntab", _i), ("coef_z", _p), ("coef_v", _p), ("ncoef", _i),
                ("dn_cell", _p), ("dz_interp0", _d), ("delta_z", _d), ("dg_z", _p), ("dg_v", _p), ("ndg", _i),
                ("dgrowth0", _d), ("H0", _d), ("h", _d), ("z1", _d), ("z2", _d), ("z3", _d), ("sigma_p", _d * 3),
                ("norm", _d), ("density_max", _d), ("z_min", _d), ("z_max", _d),
                ("ra0", _d), ("dec0", _d), ("dra", _d), ("ddec", _d),
                ("cr0", _d), ("sr0", _d), ("cd0", _d), ("sd0", _d),
                ("u1", _p), ("u2", _p), ("ux", _p), ("uy", _p), ("uz", _p), ("seed", C.c_uint64)]


class QsoSetup(object):
    """Geometry, n(z) and normalisation of one slice (draw_qso.py:196-360).  sigma_p: std of the three lognormal
    boxes of this slice (draw_qso.py:191-193)."""

    def __init__(self, NXs, NY, NZ, NX_full, dcell, i_slice, nslice, ra0, dec0, dra, ddec, zmin, zmax, sigma_p, dmax=3,
                 rho_sum=RHO_SUM):
        """rho_sum: the reference hard-codes the nominal slice's sum of ptot (constant.rho_sum, draw_qso.py:235,345);
        a caller whose slab is not a nominal slice passes scaled_rho_sum(cells) to keep the quasar density."""
        h = constant.h
        cs = cosmo_mod.cosmo(constant.omega_M_0, Ok=constant.omega_k_0, H0=100 * h)
        self.cosmo = cs
        self.NXs, self.NY, self.NZ, self.NX_full = int(NXs), int(NY), int(NZ), int(NX_full)
        self.i_slice, self.nslice = int(i_slice), int(nslice)
        DX = DY = DZ = float(dcell)
        self.dcell = DX
        self.ra0, self.dec0, self.dra, self.ddec = float(ra0), float(dec0), float(dra), float(ddec)
        LX_full, LY, LZ = NX_full * DX, NY * DY, NZ * DZ
        R0 = h * cs.r_comoving(constant.z0)
        self.x_axis = (np.arange(NXs) + 0.5) * DX + (2 * i_slice - nslice) * LX_full / (2 * nslice)
        self.y_axis = (np.arange(NY) + 0.5) * DY - LY / 2
        self.z_axis = (np.arange(NZ) + 0.5) * DZ + R0 - LZ / 2
        self.zb = (constant.z_QSO_bias_1, constant.z_QSO_bias_2, constant.z_QSO_bias_3)
        self.coef_z, self.coef_v = lognormal_coef()
        Rmin, Rmax, tanx_max, tany_max = cosmo_mod.box_limit(LX_full, LY, LZ, R0, dmax * DX)
        if zmin > 0:
            z_min = max(float(cs.r_2_z(Rmin / h)), zmin)
            Rmin = cs.r_comoving(z_min) * h
        else:
            z_min = float(cs.r_2_z(Rmin / h))
        if zmax > 0:
            z_max = min(float(cs.r_2_z(Rmax / h)), zmax)
            Rmax = cs.r_comoving(z_max) * h
        else:
            z_max = float(cs.r_2_z(Rmax / h))
        self.z_min, self.z_max = z_min, z_max
        if dra > 0 and ddec > 0:
            surface = (np.sin(np.radians(dec0 + ddec)) - np.sin(np.radians(dec0 - ddec))) * (2 * np.radians(dra))
        else:
            surface = 4 * np.arctan(tanx_max) * np.arctan(tany_max)
        nQSOexp = N_QSO_EXP * (surface * (180 / np.pi) ** 2)
        nQSOexp *= QSO_NZ_ADHOC
        nQSOexp /= nslice
        self.nQSOexp = nQSOexp
        volFrac = (surface / 3) * (Rmax ** 3 - Rmin ** 3) / LX_full / LY / LZ
        nz = tables.nz_qso_desi()
        zc = (nz[:, 0] + nz[:, 1]) / 2
        if cs.r_2_z((R0 + LZ / 2) / h) > nz[-1, 1]:
            raise ValueError("box reaches beyond the tabulated dN/dz range")          # draw_qso.py:296-298 exits
        self.dz_interp = np.linspace(zc[0], zc[-1], 2 * NZ)
        dn_cell = _interp(zc, nz[:, 2], self.dz_interp) * DZ / cs.dist_hubble(self.dz_interp)
        dn_cell = dn_cell * DX * DY / cs.r_comoving(self.dz_interp) ** 2
        self.sigma_p = tuple(float(s) for s in sigma_p)
        m = (self.dz_interp > z_min) & (self.dz_interp < z_max)
        mean_rho = dn_cell[m] / self.cond1_correction(self.dz_interp[m])
        self.density_max = float(np.max(mean_rho))
        norm = nQSOexp / rho_sum
        norm *= self.density_max / np.mean(mean_rho)
        norm /= volFrac
        if z_max > 2.1:
            n_all = dn_cell[(self.dz_interp > z_min) * (self.dz_interp < z_max)].sum()
            n_21 = dn_cell[(self.dz_interp > 2.1) * (self.dz_interp < z_max)].sum()
            norm /= n_21 / n_all
        self.norm = float(norm)
        self.dn_cell = np.append(dn_cell, np.zeros(10 * NZ))
        self.dg_z, self.dg_v, om = tables.dgrowth()
        if om != constant.omega_M_0:
            raise ValueError("Omega_M_0 ({}) != OM of dgrowth.fits ({})".format(constant.omega_M_0, om))
        self.dgrowth0 = float(_interp(self.dg_z, self.dg_v, 0.0))

    def cond1_correction(self, z):
        """z dependence of <ptot> (draw_qso.py:315-325)."""
        z1, z2, z3 = self.zb
        s1, s2, s3 = self.sigma_p
        c = _interp(self.coef_z, self.coef_v, z)
        g = lambda zb, s: np.exp((qso_a_of_z(z, zb) * s) ** 2 / 2)      # noqa: E731
        return (c * (g(z1, s1) * (z2 - z) / (z2 - z1) + g(z2, s2) * (z - z1) / (z2 - z1))
                + (1 - c) * (g(z2, s2) * (z3 - z) / (z3 - z2) + g(z3, s3) * (z - z2) / (z3 - z2)))


def scaled_rho_sum(ncells):
    """constant.rho_sum rescaled from the nominal slice to a slab of `ncells` cells (<ptot> per cell unchanged)."""
    return RHO_SUM * (float(ncells) / NOMINAL_SLICE_CELLS)


def legacy_uniforms(seed, NXs, NY, NZ):
    """The reference's five draws per z plane (draw_qso.py:425-445) from the legacy NumPy stream seeded `seed`
    (= -seed + i_slice, draw_qso.py:154), as raw uniforms: u1, u2, uz [NZ, NXs, NY], ux [NZ, NXs], uy [NZ, NY].
    Returns ((u1, u2, ux, uy, uz), RandomState) -- MJD and FIBERID are drawn from the same stream afterwards."""
    rs = np.random.RandomState(seed)
    u1, u2, uz = (np.empty((NZ, NXs, NY)) for _ in range(3))
    ux, uy = np.empty((NZ, NXs)), np.empty((NZ, NY))
    for mz in range(NZ):
        u1[mz] = rs.random_sample((NXs, NY))
        u2[mz] = rs.random_sample((NXs, NY))
        ux[mz] = rs.random_sample(NXs)
        uy[mz] = rs.random_sample(NY)
        uz[mz] = rs.random_sample((NXs, NY))
    return (u1, u2, ux, uy, uz), rs


class QsoDrawer(object):
    """Runs smk_draw_qso for one slab.  `ctx` is the smk_ctx handle of a BoxSynth (stream / device)."""

    def __init__(self, box_synth):
        if not torch.cuda.is_available():
            raise _lib.SmkError("QsoDrawer needs a CUDA device (no CPU fallback)")
        self.bs = box_synth
        self.lib = _lib.lib()
        self.device = box_synth.device

    def _dev(self, a):
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=self.device)

    def _prepared(self, st):
        """Device copies of the set-up's tables and the scalar half of smk_qso_params (cached per set-up)."""
        if getattr(st, "_prepared", None) is not None and st._prepared[2] == self.device:
            return st._prepared[0], st._prepared[1]
        p = QsoParams()
        keep = []

        def put(name, arr):
            t = self._dev(arr)
            keep.append(t)
            setattr(p, name, t.data_ptr())
        p.nxs, p.ny, p.nz, p.nx_full = st.NXs, st.NY, st.NZ, st.NX_full
        put("x_axis", st.x_axis), put("y_axis", st.y_axis), put("z_axis", st.z_axis)
        p.dx = p.dy = p.dz = st.dcell
        put("chi", st.cosmo.chi), put("zt", st.cosmo.z)
        p.ntab = len(st.cosmo.z)
        put("coef_z", st.coef_z), put("coef_v", st.coef_v)
        p.ncoef = len(st.coef_z)
        put("dn_cell", st.dn_cell)
        p.dz_interp0, p.delta_z = float(st.dz_interp[0]), float(st.dz_interp[1] - st.dz_interp[0])
        put("dg_z", st.dg_z), put("dg_v", st.dg_v)
        p.ndg = len(st.dg_z)
        p.dgrowth0, p.H0, p.h = st.dgrowth0, constant.H0, constant.h
        p.z1, p.z2, p.z3 = st.zb
        for k in range(3):
            p.sigma_p[k] = st.sigma_p[k]
        p.norm, p.density_max, p.z_min, p.z_max = st.norm, st.density_max, st.z_min, st.z_max
        p.ra0, p.dec0, p.dra, p.ddec = st.ra0, st.dec0, st.dra, st.ddec
        r0, d0 = np.radians(st.ra0), np.radians(st.dec0)
        p.cr0, p.sr0, p.cd0, p.sd0 = float(np.cos(r0)), float(np.sin(r0)), float(np.cos(d0)), float(np.sin(d0))
        st._prepared = (p, keep, self.device)
        return p, keep

    def draw(self, setup, boxln, velo=None, ix0=0, uniforms=None, seed=0, capacity=None, chunk=0, rs=None, pmf=True):
        """boxln: three device float32 [NXs, NY, NZ] tensors (boxln_1..3 of this slab, NOT exponentiated);
        velo: (vx, vy, vz) or None (rsd off).  Returns the QSO-<i>-<N>.fits columns as a dict of numpy arrays.
        pmf=False leaves out the PLATE-MJD-FIBERID strings (a Python loop over the quasars: 10 ms per 15 000 of them,
        more than the kernel); write_qso_file() builds them when they are missing."""
        st = setup
        dev = self.device
        for t in tuple(boxln) + tuple(velo or ()):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
            assert tuple(t.shape) == (st.NXs, st.NY, st.NZ), t.shape
        p, keep = self._prepared(st)
        p.ix0 = int(ix0)
        for k in range(3):
            p.boxln[k] = boxln[k].data_ptr()
            p.velo[k] = velo[k].data_ptr() if velo is not None else None
        p.rsd = int(velo is not None)
        for name in ("u1", "u2", "ux", "uy", "uz"):
            setattr(p, name, None)
        keep = list(keep)

        def put(name, arr):
            t = self._dev(arr)
            keep.append(t)
            setattr(p, name, t.data_ptr())
        if uniforms is not None:
            for name, u in zip(("u1", "u2", "ux", "uy", "uz"), uniforms):
                put(name, u)
        p.seed = int(seed) & (2 ** 64 - 1)
        if capacity is None:       # ~ norm * <ptot> of the cells pass cond1; a fraction of those survives
            capacity = int(1024 + 1e-3 * st.NXs * st.NY * st.NZ)
        while True:
            counters = torch.zeros(2, dtype=torch.int32, device=dev)
            records = torch.empty((capacity, 8), dtype=torch.float64, device=dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.bs.stream)
            _lib.check(self.lib.smk_draw_qso(self.bs.h, C.byref(p), C.c_void_p(counters.data_ptr()),
                                             C.c_void_p(records.data_ptr()), capacity))
            e1.record(self.bs.stream)
            n, nn = (int(v) for v in counters.cpu())
            self.last_kernel_ms = e0.elapsed_time(e1)
            if n <= capacity:
                break
            capacity = n + 64          # the record buffer was too small: run again with room for every quasar
        rec = records[:n].cpu().numpy()
        rec = rec[np.argsort(rec[:, 0], kind="stable")]        # (plane, ix, iy): the order of np.where in the plane loop
        key = rec[:, 0].astype(np.int64)
        cells = np.stack([key // (st.NXs * st.NY), (key // st.NY) % st.NXs, key % st.NY], axis=1)
        thing_id = (chunk * 1e9 + st.i_slice * 1e6 + np.arange(n) + 1).astype(int)               # draw_qso.py:495
        rs = rs if rs is not None else np.random.RandomState(int(seed) % (2 ** 32))
        mjd = rs.randint(51608, high=57521, size=n)
        fiberid = rs.randint(1, high=1001, size=n)
        cat = {"Z_QSO_NO_RSD": np.float32(rec[:, 1]), "Z_QSO_RSD": np.float32(rec[:, 2]), "RA": np.float32(rec[:, 3]),
               "DEC": np.float32(rec[:, 4]), "HDU": np.int32(np.ones(n) * st.i_slice), "THING_ID": thing_id,
               "PLATE": thing_id, "MJD": np.int32(mjd), "FIBERID": np.int32(fiberid),
               "XX": np.float32(rec[:, 5]), "YY": np.float32(rec[:, 6]), "ZZ": np.float32(rec[:, 7]),
               "cells": cells, "nn_cond1": nn, "f64": rec}
        if pmf:
            cat["PMF"] = pmf_strings(cat)
        return cat


def pmf_strings(cat):
    """PLATE-MJD-FIBERID of draw_qso.py:500-503 as fixed-width byte strings."""
    return np.array(["%d-%d-%d" % (t, m, f) for t, m, f in zip(cat["PLATE"], cat["MJD"], cat["FIBERID"])], dtype="S21")


def box_sigma(box):
    """np.std of a device box (draw_qso.py:191-193 compute it on the float32 slice)."""
    b = box.double()
    m = b.mean()
    return float(torch.sqrt(((b - m) ** 2).mean()).item())


COLUMNS = ("Z_QSO_NO_RSD", "Z_QSO_RSD", "RA", "DEC", "HDU", "THING_ID", "PLATE", "MJD", "FIBERID", "PMF", "XX", "YY",
           "ZZ")


def write_qso_file(path, cat, seed, ra0, dec0):
    """QSO-<i>-<N>.fits in the layout of draw_qso.py:523-563."""
    from . import fitsio_lite as fitsio
    if "PMF" not in cat:
        cat = dict(cat, PMF=pmf_strings(cat))
    f = fitsio.FITS(path, "rw", clobber=True)
    f.write([cat[c] for c in COLUMNS], names=list(COLUMNS),
            header=[{"name": "seed", "value": int(seed), "comment": "seed used to generate randoms"},
                    {"name": "ra0", "value": float(ra0), "comment": "right ascension of the box center"},
                    {"name": "dec0", "value": float(dec0), "comment": "declination of the box center"}])
    f.close()
