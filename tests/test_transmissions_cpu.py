"""DESI transmission writer (host I/O): bin/make_transmissions.py of this repo writes the
files the unmodified reference script wrote for the same spectra_merged inputs (tests/golden/ref_trans.npz, made by
tests/golden/run_reference_shimmed.py trans), and the direct-from-rows writer produces the same files."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "..", "bin")


@pytest.fixture(scope="module")
def inputs(golden_small):
    from saclaymocks_b200.healpix import radec2pix
    g = golden_small
    ids = g["merged_zfix_THING_ID"]
    order = {t: i for i, t in enumerate(g["qso_THING_ID"])}
    rows = np.array([order[t] for t in ids])
    d = {k: g["qso_" + k][rows] for k in ("RA", "DEC", "Z_QSO_NO_RSD", "Z_QSO_RSD", "HDU")}
    d.update(ids=ids, pix=radec2pix(16, d["RA"], d["DEC"], nest=True), lam=g["merged_zfix_LAMBDA"],
             flux=g["merged_zfix_FLUX"])
    return d


def check_against_golden(outd, ref):
    from saclaymocks_b200 import fitsio_lite as fitsio
    keys = sorted(k[:-5] for k in ref if k.endswith("_path"))
    got = sorted(os.path.relpath(f, outd) for f in glob.glob(outd + "/*/*/transmission-*.fits.gz"))
    assert got == sorted(str(ref[k + "_path"]) for k in keys)
    for k in keys:
        ff = fitsio.FITS(os.path.join(outd, str(ref[k + "_path"])))
        assert [h.get_extname() for h in ff] == list(ref[k + "_hdus"])
        hd, md = ff["METADATA"].read_header(), ff["METADATA"].read()
        for name in ("HPXNSIDE", "HPXNEST", "OL", "OM", "OK", "H0", "HPXPIXEL", "NSIDE"):
            assert hd[name] == ref[k + "_hdr_" + name], (k, name)
        o = np.argsort(md["MOCKID"])
        for c in ("RA", "DEC", "Z_noRSD", "Z", "MOCKID"):
            assert md[c].dtype == ref[k + "_" + c].dtype and np.array_equal(md[c][o], ref[k + "_" + c]), (k, c)
        assert np.array_equal(ff["WAVELENGTH"].read(), ref[k + "_WAVELENGTH"])
        t = ff["TRANSMISSION"].read()
        assert t.dtype == np.float32
        assert np.array_equal(t[o].astype("f8").sum(axis=1), ref[k + "_TRANSMISSION_sum"])


def test_cli_matches_reference_files(tmp_path, inputs):
    from saclaymocks_b200 import fitsio_lite as fitsio
    ref = dict(np.load(os.path.join(HERE, "golden", "ref_trans.npz")))
    d = inputs
    ind, outd = str(tmp_path / "in"), str(tmp_path / "out")
    os.makedirs(ind + "/chunk_1/spectra_merged")
    for p in np.unique(d["pix"]):
        for h in np.unique(d["HDU"][d["pix"] == p]):
            m = np.where((d["pix"] == p) & (d["HDU"] == h))[0]
            f = fitsio.FITS(ind + "/chunk_1/spectra_merged/spectra_merged-{}-{}.fits.gz".format(p, h), "rw", clobber=True)
            f.write([d["RA"][m], d["DEC"][m], d["Z_QSO_NO_RSD"][m], d["Z_QSO_RSD"][m], d["HDU"][m], d["ids"][m]],
                    names=["RA", "DEC", "Z_noRSD", "Z", "HDU", "THING_ID"], extname="METADATA")
            f.write(d["lam"], extname="LAMBDA")
            f.write(d["flux"][m], extname="FLUX")
            f.close()
    r = subprocess.run([sys.executable, os.path.join(BIN, "make_transmissions.py"), "-inDir", ind, "-outDir", outd, "-job",
                        "0", "-ncpu", "1", "-nside", "16", "-nest", "True"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "%d Sorted spectra written" % len(glob.glob(ind + "/*/spectra_merged/*")) in r.stdout
    check_against_golden(outd, ref)


def test_direct_writer_from_rows(tmp_path, inputs):
    from saclaymocks_b200 import transmissions
    ref = dict(np.load(os.path.join(HERE, "golden", "ref_trans.npz")))
    d = inputs
    outd = str(tmp_path / "out")
    files = transmissions.from_rows(outd, d["RA"], d["DEC"], d["Z_QSO_NO_RSD"], d["Z_QSO_RSD"], d["ids"], d["lam"],
                                    d["flux"])
    assert len(files) == len(np.unique(d["pix"]))
    check_against_golden(outd, ref)
