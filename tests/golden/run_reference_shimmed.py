#!/usr/bin/env python
"""Run the UNMODIFIED reference scripts (from /root/reference/bin) on small boxes and store
their outputs as golden fixtures under tests/golden/.

Build-container only (the GPU box has no /root/reference):
    python tests/golden/run_reference_shimmed.py [small|ref32|all]

The reference cannot be imported as-is here: the wheels pyfftw, fitsio,
healpy, matplotlib, h5py are absent and scipy 1.18 dropped the numpy aliases (`sp.array`...).
This script supplies
  * stand-in modules under tests/golden/ref_shims/ (fitsio -> saclaymocks_b200.fitsio_lite,
    pyfftw -> scipy.fft/pocketfft with FFTW's unnormalised conventions, healpy.ang2pix ->
    saclaymocks_b200.healpix, empty matplotlib / h5py),
  * the removed aliases `scipy.<numpy function>`, `scipy.random`, a `numpy.array` that
    falls back to dtype=object for ragged lists (NumPy < 1.24 behaviour relied upon by
    bin/make_spectra.py:536-542 and bin/merge_spectra.py:213-219), and a `numpy.linspace`
    that accepts the float `num` produced by the py2-era `nr = nk/2` (powerspectrum.py:197),
and then executes interpolate_pk.py, merge_pk.py, make_boxes.py, make_spectra.py and
merge_spectra.py through runpy with the reference's own CLI.  Everything numerical is the
reference's code; only the third-party FFT backend differs (pocketfft instead of FFTW).
QSO catalogues are synthetic (draw_qso.py is outside the hot path and needs a missing blob).
"""
import glob
import os
import runpy
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"


def install_shims():
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(HERE, "ref_shims"))
    sys.path.insert(0, os.path.join(REF, "py"))
    os.environ["SACLAYMOCKS_BASE"] = REF
    import scipy
    for name in dir(np):          # scipy < 1.x re-exported the whole numpy namespace (sp.array, sp.pi, ...)
        if name.startswith("_") or name in scipy.__dict__:
            continue
        try:
            import importlib
            importlib.import_module("scipy." + name)       # a genuine scipy submodule wins (fft, linalg, ...)
        except Exception:
            setattr(scipy, name, getattr(np, name))
    if not hasattr(scipy, "random"):
        scipy.random = np.random
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid
    _linspace = np.linspace

    def linspace_py2(start, stop, num=50, *a, **k):
        # py2-era `nr = nk/2` (py/SaclayMocks/powerspectrum.py:197) reaches linspace as a float
        return _linspace(start, stop, int(num), *a, **k)
    np.linspace = linspace_py2
    import numba    # numba must register its np.array overload before np.array is wrapped below
    numba.njit(lambda x: np.array([x, x]).sum())(1.0)
    _array = np.array

    def ragged_array(obj, *a, **k):
        try:
            return _array(obj, *a, **k)
        except ValueError:
            if a or k:
                raise
            out = np.empty(len(obj), dtype=object)
            for i, o in enumerate(obj):
                out[i] = o
            return out
    np.array = ragged_array
    scipy.array = _array      # numba-compiled kernels resolve sp.array to the real builtin


def run(script, argv):
    old = sys.argv
    sys.argv = [script] + [str(a) for a in argv]
    print(">>>", os.path.basename(script), " ".join(sys.argv[1:]), flush=True)
    try:
        runpy.run_path(os.path.join(REF, "bin", script), run_name="__main__")
    except SystemExit as e:
        if e.code not in (None, 0):
            raise
    finally:
        sys.argv = old


def synthetic_qsos(qsodir, nslice, nq, half_angle_deg, ra0, dec0, seed, zlo=1.9, zhi=3.55):
    """QSO-<i>-<nslice>.fits in the layout of bin/draw_qso.py:523-563."""
    from saclaymocks_b200 import fitsio_lite as fitsio
    rng = np.random.default_rng(seed)
    ra = (ra0 + rng.uniform(-half_angle_deg, half_angle_deg, nq)).astype("f4")
    dec = (dec0 + rng.uniform(-half_angle_deg, half_angle_deg, nq)).astype("f4")
    z = rng.uniform(zlo, zhi, nq).astype("f4")
    zrsd = (z + rng.normal(0, 0.003, nq)).astype("f4")
    order = np.argsort(ra)          # file i holds the i-th ra band, like x-slabs do
    per = int(np.ceil(nq / nslice))
    names = ["Z_QSO_NO_RSD", "Z_QSO_RSD", "RA", "DEC", "HDU", "THING_ID", "PLATE", "MJD", "FIBERID",
             "PMF", "XX", "YY", "ZZ"]
    for i in range(nslice):
        idx = order[i * per:(i + 1) * per]
        n = len(idx)
        tid = (1 * 10**9 + i * 10**6 + np.arange(n) + 1).astype("i8")
        pmf = np.array(["{}-{}-{}".format(t, 0, 0) for t in tid], dtype="S21")
        cols = [z[idx], zrsd[idx], ra[idx], dec[idx], np.full(n, i, "i4"), tid, tid.copy(),
                np.zeros(n, "i4"), np.zeros(n, "i4"), pmf, np.zeros(n, "f4"), np.zeros(n, "f4"), np.zeros(n, "f4")]
        f = fitsio.FITS(qsodir + "/QSO-{}-{}.fits".format(i, nslice), "rw", clobber=True)
        f.write(cols, names=names, header=[{"name": "seed", "value": seed}, {"name": "ra0", "value": ra0},
                                          {"name": "dec0", "value": dec0}], extname="QSO")
        f.close()


def pipeline(tag, NX, NY, NZ, dcell, nslice, nq, half_angle, seed=42, keep_full=True, stride=1):
    from saclaymocks_b200 import fitsio_lite as fitsio
    from saclaymocks_b200 import p1dmiss
    tmp = tempfile.mkdtemp(prefix="smk_ref_" + tag)
    d = {k: os.path.join(tmp, k) for k in ("pk", "boxes", "qso", "spectra", "merged_zfix", "merged_z")}
    for v in d.values():
        os.makedirs(v)
    ra0, dec0 = 190.0, 0.0
    dims = ["-NX", NX, "-NY", NY, "-NZ", NZ]
    run("interpolate_pk.py", dims + ["-pixel", dcell, "-i", 0, "-N", 1, "-outDir", d["pk"]])
    run("merge_pk.py", dims + ["-inDir", d["pk"], "-outDir", d["pk"], "-N", 1])
    run("make_boxes.py", dims + ["-pixel", dcell, "-nHDU", nslice, "-ncpu", 2, "-PkDir", d["pk"], "-seed", seed,
                                 "-rsd", "True", "-outDir", d["boxes"]])
    synthetic_qsos(d["qso"], nslice, nq, half_angle, ra0, dec0, seed)
    for i in range(nslice):
        run("make_spectra.py", ["-QSOfile", d["qso"] + "/QSO-", "-boxdir", d["boxes"], "-outDir", d["spectra"],
                                "-i", i, "-N", nslice, "-zmin", 1.8, "-zmax", 3.6, "-rsd", "True", "-dla", "True"])
    p1dfile = p1dmiss.build_pkmiss_interp(os.path.join(tmp, "pkmiss_standin.fits"))
    for i in range(nslice):
        common = ["-inDir", d["spectra"], "-i", i, "-p1dfile", p1dfile, "-seed", seed, "-rsd", "True",
                  "-dla", "True", "--store-g", "True", "-ncpu", 1, "-bb", -1]
        run("merge_spectra.py", common + ["-outDir", d["merged_zfix"], "-zfix", 2.4])
        run("merge_spectra.py", common + ["-outDir", d["merged_z"]])

    out = {"NX": NX, "NY": NY, "NZ": NZ, "dcell": dcell, "nslice": nslice, "seed": seed,
           "ra0": ra0, "dec0": dec0, "stride": stride}
    pname = "/P%d.fits" % NX if (NX == NY == NZ) else "/P%d-%d-%d.fits" % (NX, NY, NZ)
    for ext in ("Pln1", "Pln2", "Pln3", "P0"):
        w = fitsio.read(d["pk"] + pname, ext=ext)
        out["W_" + ext] = w if keep_full else w.ravel()[::stride]
    boxk = np.load(d["boxes"] + "/boxk.npy")      # = FFT(noise) * P0 at this point (make_boxes.py:289-291)
    out["boxkP0"] = boxk if keep_full else boxk.ravel()[::stride]
    nfiles = {"boxln_1": nslice, "boxln_2": nslice, "boxln_3": nslice, "vx": nslice, "vy": nslice, "vz": nslice,
              "box": NX, "eta_xx": NX, "eta_yy": NX, "eta_zz": NX, "eta_xy": NX, "eta_xz": NX, "eta_yz": NX}
    for name, n in nfiles.items():
        parts = [fitsio.read(d["boxes"] + "/%s-%d.fits" % (name, i)) for i in range(n)]
        box = np.concatenate(parts)
        h = fitsio.read_header(d["boxes"] + "/%s-0.fits" % name)
        out["sigma_" + name] = np.float32(h["sigma"])
        out["box_" + name] = box if keep_full else box.ravel()[::stride]
        out["sum_" + name] = np.float64(box.astype("f8").sum())
        out["sumsq_" + name] = np.float64((box.astype("f8") ** 2).sum())
    # QSO catalogue
    q = np.concatenate([fitsio.read(f, ext=1) for f in sorted(glob.glob(d["qso"] + "/QSO-*.fits"))])
    for c in ("RA", "DEC", "Z_QSO_NO_RSD", "Z_QSO_RSD", "THING_ID", "HDU"):
        out["qso_" + c] = q[c]
    # per-slab spectra pieces (make_spectra output)
    for f in sorted(glob.glob(d["spectra"] + "/spectra-*.fits.gz")):
        key = os.path.basename(f).split(".")[0].replace("-", "_")
        ff = fitsio.FITS(f)
        out[key + "_THING_ID"] = ff["METADATA"].read()["THING_ID"]
        out[key + "_Npixel"] = ff["METADATA"].read_header()["Npixel"]
        for ext in ("LAMBDA", "DELTA_L", "ETA_PAR", "VELO_PAR", "REDSHIFT"):
            out[key + "_" + ext] = ff[ext].read()
    # merged spectra
    for mode in ("merged_zfix", "merged_z"):
        ids, flux, dl, ep, vp, ds = [], [], [], [], [], []
        for f in sorted(glob.glob(d[mode] + "/spectra_merged-*.fits.gz")):
            ff = fitsio.FITS(f)
            ids.append(ff["METADATA"].read()["THING_ID"])
            flux.append(ff["FLUX"].read())
            dl.append(ff["DELTA_L"].read())
            ep.append(ff["ETA_PAR"].read())
            vp.append(ff["VELO_PAR"].read())
            ds.append(ff["DELTA_S"].read())
            out[mode + "_LAMBDA"] = ff["LAMBDA"].read()
            out[mode + "_GROWTHF"] = ff["GROWTHF"].read()
            out[mode + "_Z"] = ff["Z"].read()
            out[mode + "_pixfile_" + os.path.basename(f).split(".")[0]] = ff["METADATA"].read()["THING_ID"]
        out[mode + "_THING_ID"] = np.concatenate(ids)
        out[mode + "_FLUX"] = np.concatenate(flux)
        out[mode + "_DELTA_L"] = np.concatenate(dl)
        out[mode + "_ETA_PAR"] = np.concatenate(ep)
        out[mode + "_VELO_PAR"] = np.concatenate(vp)
        out[mode + "_DELTA_S"] = np.concatenate(ds)
    dst = os.path.join(HERE, "ref_%s.npz" % tag)
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KiB")
    shutil.rmtree(tmp)


def qso_case(tag, NX, NY, NZ, dcell, nslice, dra, ddec, seed=42, zfix=None):
    """make_boxes.py -> draw_qso.py (one run per slice, -desi False: the DESI footprint map is a missing blob)
    on a small box; stores the six input boxes of draw_qso and every column of the QSO-<i>-<nslice>.fits tables."""
    from saclaymocks_b200 import fitsio_lite as fitsio
    tmp = tempfile.mkdtemp(prefix="smk_refqso_" + tag)
    d = {k: os.path.join(tmp, k) for k in ("pk", "boxes", "qso")}
    for v in d.values():
        os.makedirs(v)
    ra0, dec0 = 190.0, 0.0
    dims = ["-NX", NX, "-NY", NY, "-NZ", NZ]
    run("interpolate_pk.py", dims + ["-pixel", dcell, "-i", 0, "-N", 1, "-outDir", d["pk"]])
    run("merge_pk.py", dims + ["-inDir", d["pk"], "-outDir", d["pk"], "-N", 1])
    run("make_boxes.py", dims + ["-pixel", dcell, "-nHDU", nslice, "-ncpu", 2, "-PkDir", d["pk"], "-seed", seed,
                                 "-rsd", "True", "-outDir", d["boxes"]])
    out = {"NX": NX, "NY": NY, "NZ": NZ, "dcell": dcell, "nslice": nslice, "seed": seed, "ra0": ra0, "dec0": dec0,
           "dra": dra, "ddec": ddec, "chunk": 1, "zmin": 1.8, "zmax": 3.6, "zfix": -1.0 if zfix is None else zfix}
    for name in ("boxln_1", "boxln_2", "boxln_3", "vx", "vy", "vz"):
        out["box_" + name] = np.concatenate([fitsio.read(d["boxes"] + "/%s-%d.fits" % (name, i)) for i in range(nslice)])
    for i in range(nslice):
        a = ["-indir", d["boxes"], "-outpath", d["qso"], "-i", i, "-Nslice", nslice, "-chunk", 1, "-ra0", ra0,
             "-dec0", dec0, "-dra", dra, "-ddec", ddec, "-zmin", 1.8, "-zmax", 3.6, "-desi", "False", "-seed", seed,
             "-rsd", "True"]
        if zfix is not None:
            a += ["-zfix", zfix]
        run("draw_qso.py", a)
        f = fitsio.FITS(d["qso"] + "/QSO-%d-%d.fits" % (i, nslice))
        t = f[1].read()
        hd = f[1].read_header()
        out["qso%d_seed" % i] = hd["seed"]
        for c in ("Z_QSO_NO_RSD", "Z_QSO_RSD", "RA", "DEC", "HDU", "THING_ID", "PLATE", "MJD", "FIBERID", "PMF",
                  "XX", "YY", "ZZ"):
            out["qso%d_%s" % (i, c)] = t[c]
        print("slice", i, "nqso", len(t))
    dst = os.path.join(HERE, "ref_%s.npz" % tag)
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KiB")
    shutil.rmtree(tmp)


def trans_case():
    """make_transmissions.py on spectra_merged files rebuilt from ref_small.npz (FLUX rows of the -zfix run): stores,
    per transmission file, the header keys, MOCKIDs (sorted: the reference concatenates in glob order) and HDU names."""
    from saclaymocks_b200 import fitsio_lite as fitsio
    from saclaymocks_b200.healpix import radec2pix
    g = dict(np.load(os.path.join(HERE, "ref_small.npz")))
    tmp = tempfile.mkdtemp(prefix="smk_reftrans")
    ind, outd = tmp + "/in", tmp + "/out"
    ids = g["merged_zfix_THING_ID"]
    order = {t: i for i, t in enumerate(g["qso_THING_ID"])}
    rows = np.array([order[t] for t in ids])
    ra, dec, zn, zr, hdu = (g["qso_" + k][rows] for k in ("RA", "DEC", "Z_QSO_NO_RSD", "Z_QSO_RSD", "HDU"))
    pix = radec2pix(16, ra, dec, nest=True)
    os.makedirs(ind + "/chunk_1/spectra_merged")
    for p in np.unique(pix):
        os.makedirs(outd + "/{}/{}".format(p // 100, p), exist_ok=True)
        for h in np.unique(hdu[pix == p]):
            m = np.where((pix == p) & (hdu == h))[0]
            f = fitsio.FITS(ind + "/chunk_1/spectra_merged/spectra_merged-{}-{}.fits.gz".format(p, h), "rw", clobber=True)
            f.write([ra[m], dec[m], zn[m], zr[m], hdu[m], ids[m]], names=["RA", "DEC", "Z_noRSD", "Z", "HDU", "THING_ID"],
                    extname="METADATA")
            f.write(g["merged_zfix_LAMBDA"], extname="LAMBDA")
            f.write(g["merged_zfix_FLUX"][m], extname="FLUX")
            f.close()
    run("make_transmissions.py", ["-inDir", ind, "-outDir", outd, "-job", 0, "-ncpu", 1, "-nside", 16, "-nest", "True"])
    out = {}
    for f in sorted(glob.glob(outd + "/*/*/transmission-*.fits.gz")):
        key = os.path.basename(f).split(".")[0].replace("-", "_")
        ff = fitsio.FITS(f)
        hd = ff["METADATA"].read_header()
        md = ff["METADATA"].read()
        o = np.argsort(md["MOCKID"])
        out[key + "_path"] = os.path.relpath(f, outd)
        out[key + "_hdus"] = np.array([h.get_extname() for h in ff])
        for k in ("HPXNSIDE", "HPXNEST", "OL", "OM", "OK", "H0", "HPXPIXEL", "NSIDE"):
            out[key + "_hdr_" + k] = hd[k]
        for c in ("RA", "DEC", "Z_noRSD", "Z", "MOCKID"):
            out[key + "_" + c] = md[c][o]
        out[key + "_WAVELENGTH"] = ff["WAVELENGTH"].read()
        out[key + "_TRANSMISSION_sum"] = ff["TRANSMISSION"].read()[o].astype("f8").sum(axis=1)
    dst = os.path.join(HERE, "ref_trans.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KiB", len(out))
    shutil.rmtree(tmp)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    install_shims()
    if which in ("trans",):
        trans_case()
    if which in ("qso", "all"):
        # 32 x 32 x 192 cells of 17.52 Mpc/h (LZ = 3364 Mpc/h like the nominal box), 2 slices
        qso_case("qso", 32, 32, 192, 17.52, 2, dra=30.0, ddec=30.0)   # wide cut: ~1e3 quasars per slice
    if which in ("small", "all"):
        # 16 x 16 x 96 cells of 35.04 Mpc/h (LZ = 3364 Mpc/h like the nominal box), 4 x-slabs so that
        # some sightlines cross a slab boundary and exercise the piece merge (merge_spectra.py:282-300)
        pipeline("small", 16, 16, 96, 35.04, 4, nq=16, half_angle=2.0, keep_full=True)
    if which in ("c1",):
        # BASELINE config 1 (submit_mocks.py --box-size 256: 256 x 256 x 1536 cells of 2.19 Mpc/h, 8 slices,
        # chunk_parameters(256) window of +-3.17 deg): every box sampled with a stride, 24 sightlines.  Takes ~15 min and
        # ~10 GB of scratch; not part of "all".
        pipeline("c1", 256, 256, 1536, 2.19, 8, nq=24, half_angle=3.0, keep_full=False, stride=9973)
    if which in ("c2",):
        # BASELINE config 2, the bench workload (512 x 512 x 1536 cells of 2.19 Mpc/h, 8 slices, window of +-6.4 deg):
        # every box sampled with a stride, 32 sightlines.  ~25 GB of scratch, ~30 min; not part of "all".
        pipeline("c2", 512, 512, 1536, 2.19, 8, nq=32, half_angle=6.0, keep_full=False, stride=39989)
    if which in ("ref32", "all"):
        # the reference's own debugging box chunk_parameters(32): 32 x 32 x 1536 cells of 2.19 (4 slabs here)
        pipeline("ref32", 32, 32, 1536, 2.19, 4, nq=12, half_angle=0.3, keep_full=False, stride=97)
