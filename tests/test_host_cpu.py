"""CPU tests of the host substrate: FITS reader/writer, healpix, cosmology, P(k) machinery, catalogue set-up."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def test_fits_roundtrip(tmp_path):
    from saclaymocks_b200 import fitsio_lite as fitsio
    box = np.random.default_rng(0).standard_normal((3, 4, 5)).astype(np.float32)
    fn = str(tmp_path / "box-0.fits")
    f = fitsio.FITS(fn, "rw", clobber=True)
    f.write(box, header={"DX": 2.19, "NX": 256})
    f[0].write_key("sigma", np.float32(1.5), comment="std of the box")
    f[0].write_key("seed", np.int32(42))
    f.close()
    assert os.path.getsize(fn) % 2880 == 0
    got, h = fitsio.read(fn, header=True)
    assert np.array_equal(got, box) and got.dtype == np.float32
    assert (h["NAXIS1"], h["NAXIS2"], h["NAXIS3"]) == (5, 4, 3)       # NAXIS1 is the fastest (z) axis
    assert h["DX"] == 2.19 and h["nx"] == 256 and h["SIGMA"] == 1.5 and h["seed"] == 42
    gz = str(tmp_path / "spectra-0-1.fits.gz")
    g = fitsio.FITS(gz, "rw", clobber=True)
    cols = [np.float32([1, 2]), np.int64([10**9 + 1, 10**9 + 2]), np.array([b"a-b-c", b"d"], dtype="S21"), np.int32([3, 4])]
    g.write(cols, names=["RA", "THING_ID", "PMF", "HDU"], header=[{"name": "z0", "value": 1.7, "comment": "c"},
                                                                    {"name": "Npixel", "value": 6524}], extname="METADATA")
    g.write(np.float32([[1, 2, 3], [4, 5, -1]]), extname="LAMBDA")
    g.close()
    r = fitsio.FITS(gz)
    t = r["METADATA"].read()
    assert list(t["THING_ID"]) == [10**9 + 1, 10**9 + 2] and t["PMF"][0] == b"a-b-c" and t.dtype["RA"] == np.float32
    assert r[1].read_header()["NPIXEL"] == 6524 and r["lambda"].read()[1, 2] == -1
    with pytest.raises(KeyError):
        r["NOPE"]


def test_fits_reads_reference_style_tables():
    from saclaymocks_b200 import tables
    z, a, b, c = tables.params()
    assert len(z) == 10000 and abs(z[0] - 1.8) < 1e-12 and abs(a[0] - 3.0e-6) < 1e-12      # SURVEY 8c anchors
    i = np.argmin(np.abs(z - 2.4))
    assert abs(a[i] - 0.012764) < 2e-5 and abs(b[i] - 1.6466) < 2e-4
    zz, dd, om = tables.dgrowth()
    assert om == 0.31457 and abs(dd[0] + 0.51383754) < 1e-8
    zs, k, pk = tables.p1dmiss_tables()
    assert pk.shape == (5, 20000) and k[1] == 0.001


def test_healpix_properties():
    from saclaymocks_b200.healpix import ang2pix, radec2pix
    rng = np.random.default_rng(1)
    z = rng.uniform(-1, 1, 100000)
    phi = rng.uniform(0, 2 * np.pi, 100000)
    for nside in (1, 4, 16):
        for nest in (True, False):
            p = ang2pix(nside, np.arccos(z), phi, nest=nest)
            assert p.min() >= 0 and p.max() < 12 * nside ** 2
            c = np.bincount(p, minlength=12 * nside ** 2)
            if nside <= 4:
                assert abs(c / c.mean() - 1).max() < 0.2                    # equal-area pixels
    # nested and ring label the same pixels: a bijection between the two numberings
    pn = ang2pix(8, np.arccos(z), phi, nest=True)
    pr = ang2pix(8, np.arccos(z), phi, nest=False)
    assert len(set(zip(pn, pr))) == len(set(pn)) == len(set(pr))
    assert ang2pix(1, 0.1, 0.1, nest=True) == 0 and ang2pix(1, np.pi - 0.1, 0.1, nest=True) == 8
    assert radec2pix(16, np.array([190.0]), np.array([0.0]))[0] == ang2pix(16, np.pi / 2, np.radians(190.0), nest=True)


def test_cosmology_and_geometry_known_answers():
    from saclaymocks_b200 import cosmo, constant
    from saclaymocks_b200 import spectra as sp
    assert abs(cosmo.fgrowth(2.4, 0.31457) - 0.370199667909) < 1e-11
    c = cosmo.cosmo()
    assert abs(constant.h * c.r_comoving(constant.z0) - 3273.6836790) < 1e-6
    with pytest.raises(ValueError):
        c.r_comoving(11.0)
    g = sp.SkewerGeometry(256, 256, 1536, 2.19)
    assert g.npixeltot == 6524 and abs(g.R0 - 3273.6836790) < 1e-6
    x, y, z = cosmo.ComputeXYZ2(np.radians(np.float32(190.0)), np.radians(np.float32(0.0)), 1000.0, np.radians(190.0), 0.0)
    assert abs(x) < 1e-3 and abs(y) < 1e-3 and abs(z - 1000.0) < 1e-4     # float32 trigonometry, like the reference
    Rmin, Rmax, tx, ty = cosmo.box_limit(560.64, 560.64, 3363.84, 3273.68, 6.57)
    assert tx == ty and 0.05 < tx < 0.06 and Rmax == 3273.68 + 3363.84 / 2 - 6.57


def test_pk_host_tables_and_ppoly(golden_small):
    from saclaymocks_b200 import pk
    W = pk.weight_tables(16, 16, 96, 35.04, 4, 12)
    for k in W:
        assert np.array_equal(W[k], golden_small["W_" + k][4:12])         # kx-row slicing of interpolate_pk -i/-N
    br, co = pk.ppoly("P0")
    kk = np.linspace(0, 2.5, 4001)
    i = np.clip(np.searchsorted(br, kk, side="right") - 1, 0, co.shape[1] - 1)
    dx = kk - br[i]
    v = ((co[0, i] * dx + co[1, i]) * dx + co[2, i]) * dx + co[3, i]
    assert np.max(np.abs(v - pk.spline("P0")(kk))) < 1e-9 * np.max(pk.spline("P0")(kk))


def test_catalogue_setup_matches_oracle(golden_small):
    """qso_lines_of_sight (vectorised make_spectra.py:429-431, 467-472) against the oracle's per-quasar values."""
    from oracle import cosmology as co
    from saclaymocks_b200 import spectra as sp
    from helpers import qso_files_from_golden
    g = golden_small
    q = np.concatenate(qso_files_from_golden(g))
    geom = sp.SkewerGeometry(int(g["NX"]), int(g["NY"]), int(g["NZ"]), float(g["dcell"]))
    xyzr, nfor = sp.qso_lines_of_sight(geom, q["RA"], q["DEC"], q["Z_QSO_RSD"], float(g["ra0"]), float(g["dec0"]))
    cos = co.Cosmo()
    R_vec, lam = co.pixel_grid(cos)
    assert np.array_equal(R_vec, geom.R_vec) and np.array_equal(lam, geom.lambda_vec)
    for i in range(len(q)):
        R = co.h * cos.r_comoving(q["Z_QSO_RSD"][i])
        X, Y, Z = co.compute_xyz2(np.radians(q["RA"][i]), np.radians(q["DEC"][i]), R, np.radians(float(g["ra0"])),
                                  np.radians(float(g["dec0"])))
        assert np.allclose(xyzr[i], [X, Y, Z, R], rtol=0, atol=1e-9)
        w = np.where(lam < co.lya * (1 + q["Z_QSO_RSD"][i]))[0]
        assert nfor[i] == (0 if len(w) == 0 else w[-1] + 1)
    fg_count = sp.FGPA.forest_count.__get__(type("S", (), {"geom": geom})())(q["Z_QSO_RSD"])
    lam32 = np.float32(lam)
    for i in range(len(q)):
        rf = lam32 / (np.float32(1) + q["Z_QSO_RSD"][i])
        assert fg_count[i] == int(((rf < np.float32(co.lya)) & (rf > 0)).sum())     # merge_spectra.py:305-306


def test_reference_arm_of_bench_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "grf_cells_per_s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_chunk_parameters_match_the_reference_table():
    """saclaymocks_b200.chunks.chunk_parameters against the unmodified submit_mocks.py:611-674 (fixture written by
    tests/golden/make_ref_chunks.py): same strings, same order, same slab counts, for every box size and the stripe
    footprint; chunk 1 of the nominal box and the C1 / C2 windows of SURVEY.md section 8d as floats."""
    from saclaymocks_b200 import chunks
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_chunks.json")))
    assert len(g) == 12
    for key, v in g.items():
        cells, stripe = (int(t) for t in key.split("_"))
        ra0, dra, dec0, ddec, cid, nslice = chunks.chunk_parameters(cells, bool(stripe))
        for name, arr in (("ra0", ra0), ("dra", dra), ("dec0", dec0), ("ddec", ddec), ("chunkid", cid)):
            assert isinstance(arr, np.ndarray) and arr.dtype.kind == "U" and list(arr) == v[name], (key, name)
        assert int(nslice) == v["nslice"] and nslice.shape == ()
    assert chunks.chunk_ids(2560) == [1, 2, 3, 4, 5, 6, 7]
    assert chunks.chunk_window(2560, 1) == (125.5, 32.2413248675, 20.0, 32.2413248675)
    assert chunks.chunk_window(256, 1) == (189.982649735, 3.17, 20.0, 3.17)
    assert chunks.chunk_window(512, 6) == (202.8, 6.4, 12.8, 6.4)
    with pytest.raises(ValueError):
        chunks.chunk_parameters(100)


def test_bench_keeps_stdout_for_the_json_line():
    """bench.quiet_stdout / emit: anything written to fd 1 between the two (python prints, a child process, NCCL's
    banner) lands on stderr; stdout carries exactly the emitted line."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; fd = bench.quiet_stdout(); print('junk'); "
            "os.system('echo junk2'); bench.emit(fd, '{\"ok\": 1}')" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout == '{"ok": 1}\n'
    assert "junk" in r.stderr and "junk2" in r.stderr
