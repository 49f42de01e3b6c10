"""Pin the CPU oracle against outputs of the unmodified reference scripts
(tests/golden/ref_small.npz, ref_ref32.npz made by tests/golden/run_reference_shimmed.py)."""
import numpy as np
import pytest

from oracle import boxes as oboxes
from oracle import cosmology as co
from oracle import merge as omerge
from oracle import pk_weights
from oracle import spectra as ospectra
from helpers import qso_files_from_golden, rel_l2


def test_known_answers():
    """SURVEY.md Appendix B anchors."""
    assert abs(co.fgrowth(2.4, 0.31457) - 0.370199667909) < 1e-11
    assert abs(co.fgrowth(2.3) - 0.381076279483) < 1e-11
    cosmo = co.Cosmo()
    assert abs(co.h * cosmo.r_comoving(co.z0) - 3273.6836790) < 1e-6
    assert abs(co.h * cosmo.r_comoving(1.8) - 3374.0922817) < 1e-6
    assert abs(co.h * cosmo.r_comoving(3.6) - 4742.7198389) < 1e-6
    R, lam = co.pixel_grid(cosmo)
    assert len(R) == 6524
    assert abs(oboxes.dgrowth0() + 0.51383754) < 1e-8
    G = co.fgrowth(2.4)
    a, b, c = 0.012763985186186, 1.646616001600495, 1.673208333333331
    for (d, e), F in {(0, 0): 0.9873171290, (1, 0.2): 0.9716158422, (-1, -0.3): 0.9949037110,
                      (2.5, 0.5): 0.9070411807}.items():
        assert abs(omerge.fgpa(d, e, G, a, b, c) - F) < 1e-9
    assert omerge.fgpa(-1e6, 0., G, a, b, c) == 1.0
    box = oboxes.draw_noise(4, 4, 2, 42)
    np.testing.assert_allclose(box[0, 0:4, 0], np.float32([0.49671414, -0.1382643, 0.64768857, 1.5230298]))


@pytest.fixture(scope="module")
def small_run(golden_small):
    g = golden_small
    NX, NY, NZ, dcell = int(g["NX"]), int(g["NY"]), int(g["NZ"]), float(g["dcell"])
    W = pk_weights.weights(NX, NY, NZ, dcell)
    raw, p0, boxes, sig = oboxes.make_boxes(NX, NY, NZ, dcell, int(g["seed"]), W)
    return W, raw, p0, boxes, sig


def test_weights_match_interpolate_pk(golden_small, small_run):
    W = small_run[0]
    for k in ("Pln1", "Pln2", "Pln3", "P0"):
        np.testing.assert_array_equal(W[k], golden_small["W_" + k])


def test_boxes_match_make_boxes(golden_small, small_run):
    g = golden_small
    W, raw, p0, boxes, sig = small_run
    assert rel_l2(p0, g["boxkP0"]) < 1e-7
    for name in oboxes.PRODUCTS:
        assert rel_l2(boxes[name], g["box_" + name]) < 2e-7, name
        assert abs(sig[name] / g["sigma_" + name] - 1) < 1e-6, name


def _run_spectra(g, boxes):
    NX, NY, NZ, dcell, nslice = int(g["NX"]), int(g["NY"]), int(g["NZ"]), float(g["dcell"]), int(g["nslice"])
    geom = ospectra.Geometry(NX, NY, NZ, dcell)
    qf = qso_files_from_golden(g)
    return geom, [ospectra.make_spectra_slice(geom, boxes, qf, i, nslice, float(g["ra0"]), float(g["dec0"]))
                  for i in range(nslice)]


def _check_pieces(g, all_pieces, tol):
    nfound = 0
    for islice, pieces in enumerate(all_pieces):
        for hdu in sorted(set(p["hdu"] for p in pieces)):
            key = "spectra_%d_%d" % (islice, hdu)
            sel = [p for p in pieces if p["hdu"] == hdu]
            ids = g[key + "_THING_ID"]
            assert [p["id"] for p in sel] == list(ids)
            for r, p in enumerate(sel):
                n = len(p["lam"])
                nfound += 1
                np.testing.assert_array_equal(g[key + "_LAMBDA"][r][:n], p["lam"])
                assert np.all(g[key + "_LAMBDA"][r][n:] == -1)
                np.testing.assert_array_equal(g[key + "_REDSHIFT"][r][:n], p["redshift"])
                for ext, fld in (("DELTA_L", "delta_l"), ("ETA_PAR", "eta_par"), ("VELO_PAR", "velo_par")):
                    ref = g[key + "_" + ext][r][:n]
                    scale = max(1.0, float(np.abs(ref[ref > -1e5]).max())) if np.any(ref > -1e5) else 1.0
                    assert np.max(np.abs(ref - p[fld])) <= tol * scale, (key, ext)
    assert nfound == sum(1 for k in g if k.startswith("spectra_") and k.endswith("_THING_ID")
                         for _ in g[k])


def test_spectra_match_make_spectra_small(golden_small, small_run):
    boxes = {k: golden_small["box_" + k] for k in ospectra.FIELDS}     # identical boxes: isolate the gather
    geom, pieces = _run_spectra(golden_small, boxes)
    assert all(geom.npixeltot == int(v) for k, v in golden_small.items() if k.endswith("_Npixel"))
    _check_pieces(golden_small, pieces, 2e-7)


@pytest.mark.parametrize("mode", ["merged_zfix", "merged_z"])
def test_merge_matches_merge_spectra_small(golden_small, mode):
    g = golden_small
    boxes = {k: g["box_" + k] for k in ospectra.FIELDS}
    geom, all_pieces = _run_spectra(g, boxes)
    p1d = omerge.P1DMissing()
    flat = [p for pieces in all_pieces for p in pieces]
    merged = []
    for hdu in range(int(g["nslice"])):
        sel = [p for p in flat if p["hdu"] == hdu]
        if sel:
            merged += omerge.merge_spectra_hdu(sel, hdu, int(g["seed"]), p1d, geom.npixeltot,
                                               zfix=2.4 if mode == "merged_zfix" else None)
    ref_ids = list(g[mode + "_THING_ID"])
    got = {m["id"]: m for m in merged}
    assert sorted(got) == sorted(ref_ids)
    for r, ID in enumerate(ref_ids):
        m = got[ID]
        np.testing.assert_array_equal(m["lam"], g[mode + "_LAMBDA"])
        assert np.max(np.abs(m["delta_s"] - g[mode + "_DELTA_S"][r])) < 2e-6
        assert np.max(np.abs(m["flux"] - g[mode + "_FLUX"][r])) < 1e-6
        assert np.max(np.abs(m["delta_l"] - g[mode + "_DELTA_L"][r])) < 1e-6
        assert np.max(np.abs(m["eta_par"] - g[mode + "_ETA_PAR"][r])) < 1e-6
    np.testing.assert_allclose(got[ref_ids[0]]["growthf"], g[mode + "_GROWTHF"], rtol=1e-6)


def test_ref32_end_to_end(golden_ref32):
    """The reference's own 32 x 32 x 1536 debugging box, seed 42, from noise to FLUX."""
    _end_to_end(golden_ref32)


def test_c1_end_to_end(golden_c1):
    """BASELINE config 1 at full size (256 x 256 x 1536 cells, 8 slices, seed 42): the oracle against the unmodified
    reference scripts, from noise to FLUX (boxes compared on a strided sample of 1e4 cells each)."""
    _end_to_end(golden_c1)


def test_c2_end_to_end(golden_c2_slow):
    """BASELINE config 2, the bench workload, at full size (512 x 512 x 1536): minutes of CPU work, SMK_SLOW_TESTS=1."""
    _end_to_end(golden_c2_slow)


def _end_to_end(g):
    NX, NY, NZ, dcell, st = int(g["NX"]), int(g["NY"]), int(g["NZ"]), float(g["dcell"]), int(g["stride"])
    W = pk_weights.weights(NX, NY, NZ, dcell)
    for k in ("Pln1", "Pln2", "Pln3", "P0"):
        np.testing.assert_array_equal(W[k].ravel()[::st], g["W_" + k])
    raw, p0, boxes, sig = oboxes.make_boxes(NX, NY, NZ, dcell, int(g["seed"]), W)
    assert rel_l2(p0.ravel()[::st], g["boxkP0"]) < 1e-6
    for name in oboxes.PRODUCTS:
        assert rel_l2(boxes[name].ravel()[::st], g["box_" + name]) < 1e-6, name
        assert abs(sig[name] / g["sigma_" + name] - 1) < 1e-5, name
    geom, all_pieces = _run_spectra(g, boxes)
    _check_pieces(g, all_pieces, 5e-6)
    p1d = omerge.P1DMissing()
    flat = [p for pieces in all_pieces for p in pieces]
    for mode, zfix in (("merged_zfix", 2.4), ("merged_z", None)):
        merged = []
        for hdu in range(int(g["nslice"])):
            sel = [p for p in flat if p["hdu"] == hdu]
            if sel:
                merged += omerge.merge_spectra_hdu(sel, hdu, int(g["seed"]), p1d, geom.npixeltot, zfix=zfix)
        got = {m["id"]: m for m in merged}
        ref_ids = list(g[mode + "_THING_ID"])
        assert sorted(got) == sorted(ref_ids)
        for r, ID in enumerate(ref_ids):
            assert np.max(np.abs(got[ID]["flux"] - g[mode + "_FLUX"][r])) < 1e-5
