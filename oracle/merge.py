"""Piece merge, small-scale 1-D field and FGPA — restated from bin/merge_spectra.py:58-118
(parameter tables, seed, P1D_missing), :124-220 (piece collection), :246-339 (per-forest loop)
and py/SaclayMocks/util.py:421-433 (fgpa), :448-474 (InterpP1Dmissing), :372-398 (sigma_p1d)."""
import numpy as np
import scipy.fft as sfft

from . import cosmology as co
from saclaymocks_b200 import healpix   # host glue only: fixes the (healpix, THING_ID) iteration order


def fgpa(delta, eta_par, growthf, a, b, c):
    """util.py:421-433."""
    tau_over_a = np.exp(b * growthf * (delta + c * eta_par))
    return np.exp(-a * tau_over_a)


def sigma_p1d_row(k, pk_row, pixel=0.2, N=10000):
    """util.py:372-398 for one tabulated redshift."""
    L = N * pixel
    kj = 2 * np.pi / L * np.arange(1, N / 2)
    var_s = 2 * co.interp1d(k, pk_row, kj).sum() / L
    var_s += co.interp1d(k, pk_row, 0.0) / L
    var_s += co.interp1d(k, pk_row, np.pi / pixel) / L
    return np.sqrt(var_s)


class P1DMissing(object):
    """pkmiss_interp-format tables built from the shipped etc/p1dmiss_z*.fits (SURVEY §7 'Missing blobs').
    __call__: nearest tabulated z, then linear in k (util.py:462-474)."""

    def __init__(self, pixel=0.2):
        t = co.tables()
        self.z, self.k, self.pk = t["p1dmiss_z"], t["p1dmiss_k"], t["p1dmiss_pk"]
        self.sigma = np.array([sigma_p1d_row(self.k, self.pk[i], pixel) for i in range(len(self.z))])

    def __call__(self, redshift, k):
        iz = np.argsort(np.abs(self.z - redshift))[0]
        return co.interp1d(self.k, self.pk[iz], k)

    def sigma_s(self, z):
        return co.interp1d(self.z, self.sigma, z)


def small_scale_field(noise, n_keep, zeff, z, p1d, pixsize=0.2):
    """merge_spectra.py:310-324 given the white-noise draw `noise` (length nz)."""
    nz = len(noise)
    k_ny = np.pi / pixsize
    delta_sk = sfft.rfftn(noise)
    k = np.fft.rfftfreq(nz) * 2 * k_ny
    pmis = p1d(zeff, k)
    pmis[pmis < 0] = 0
    delta_sk *= np.sqrt(pmis / pixsize)
    delta_s = sfft.irfftn(delta_sk)
    delta_s = delta_s[0:n_keep]
    delta_s *= p1d.sigma_s(z) / p1d.sigma_s(zeff)
    return delta_s


def merge_spectra_hdu(pieces, islice, seed, p1d, npixeltot, zfix=None, rsd=True, add_noise=True,
                      aa=-1, bb=-1, cc=-1, nside=16, nest=True, pixsize=0.2, return_noise=False):
    """One merge_spectra.py process (-i islice): `pieces` are all make_spectra pieces whose QSO HDU == islice,
    in the order merge_spectra would read them (os.listdir order is arbitrary; the result does not depend on it
    because pieces are re-sorted by wavelength, merge_spectra.py:282-284).
    Returns list of dict(id, pix, flux, delta_l, eta_par, velo_par, delta_s, lam, z, growthf)."""
    t = co.tables()
    pz, pa, pb, pc = t["params_z"], t["params_a"], t["params_b"], t["params_c"]
    np.random.seed(seed + islice)                                   # merge_spectra.py:83-84
    growthf_24 = co.fgrowth(2.4, co.omega_M_0)
    IDs = np.array([p["id"] for p in pieces])
    RA = np.array([p["ra"] for p in pieces])
    DEC = np.array([p["dec"] for p in pieces])
    ZQ = np.array([p["z"] for p in pieces])
    hp = healpix.radec2pix(nside, RA, DEC, nest=nest)
    out = []
    noises = []
    for pix in np.unique(hp):
        cut = np.where(hp == pix)[0]
        for ID in np.unique(IDs[cut]):
            msk = cut[IDs[cut] == ID]
            sel = [pieces[i] for i in msk]
            wav = np.concatenate([p["lam"] for p in sel])
            order = np.argsort(wav)
            wav = wav[order]
            if zfix:
                z = zfix * np.ones_like(wav)
            else:
                z = np.concatenate([p["redshift"] for p in sel])[order]
            a = co.interp1d(pz, pa, z) if aa <= 0 else aa
            b = co.interp1d(pz, pb, z) if bb <= 0 else bb
            c = co.interp1d(pz, pc, z) if cc <= 0 else cc
            growthf = growthf_24 * (1 + 2.4) / (1 + z)
            eta = np.concatenate([p["eta_par"] for p in sel])[order]
            delta_l = np.concatenate([p["delta_l"] for p in sel])[order]
            vpar = np.concatenate([p["velo_par"] for p in sel])[order]
            delta_s = np.zeros_like(delta_l)
            delta = delta_l
            noise = None
            if add_noise:
                wav_rf = wav / (1 + ZQ[msk][0])
                mmm = np.where((wav_rf < co.lya) & (wav_rf > co.lylimit))[0]
                if len(mmm) > 0:
                    nz = 256
                    while nz < len(wav) + 50:
                        nz *= 2
                    noise = np.random.normal(size=nz)
                    noises.append(noise)
                    zeff = z[mmm].mean()
                    delta_s = small_scale_field(noise, len(wav), zeff, z, p1d, pixsize)
                    delta = delta_l + delta_s
            if rsd:
                spec = fgpa(delta, eta, growthf, a, b, c)
            else:
                spec = np.exp(-a * np.exp(b * growthf * delta))
            if len(spec) != npixeltot:
                continue
            out.append(dict(id=int(ID), pix=int(pix), flux=np.float32(spec), delta_l=np.float32(delta_l),
                            eta_par=np.float32(eta), velo_par=np.float32(vpar), delta_s=np.float32(delta_s),
                            lam=np.float32(wav), z=np.float32(z), growthf=np.float32(growthf), noise=noise))
    if return_noise:
        return out, noises
    return out
