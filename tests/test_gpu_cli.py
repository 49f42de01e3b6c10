"""Drop-in check: this repo's bin/ scripts, driven with the reference's CLI, reproduce the files the unmodified
reference scripts wrote for the same inputs (tests/golden/ref_small.npz)."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import rel_l2, write_qso_files  # noqa: E402

BIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "bin")


def run(script, *argv):
    cmd = [sys.executable, os.path.join(BIN, script)] + [str(a) for a in argv]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Error" not in r.stdout and "Traceback" not in r.stderr      # test_cor.py:141-156 greps the logs for these
    return r.stdout


def test_cli_chain_matches_reference_files(tmp_path, golden_small):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from saclaymocks_b200 import fitsio_lite as fitsio
    from saclaymocks_b200 import p1dmiss
    g = golden_small
    NX, NY, NZ, dcell, ns = int(g["NX"]), int(g["NY"]), int(g["NZ"]), float(g["dcell"]), int(g["nslice"])
    d = {k: str(tmp_path / k) for k in ("pk", "boxes", "qso", "spectra", "merged")}
    for v in d.values():
        os.makedirs(v)
    dims = ["-NX", NX, "-NY", NY, "-NZ", NZ]
    run("interpolate_pk.py", *dims, "-pixel", dcell, "-i", 0, "-N", 2, "-outDir", d["pk"])
    run("interpolate_pk.py", *dims, "-pixel", dcell, "-i", 1, "-N", 2, "-outDir", d["pk"])
    run("merge_pk.py", *dims, "-inDir", d["pk"], "-outDir", d["pk"], "-N", 2)
    P = d["pk"] + "/P%d-%d-%d.fits" % (NX, NY, NZ)
    for ext in ("Pln1", "Pln2", "Pln3", "P0"):
        assert np.array_equal(fitsio.read(P, ext=ext), g["W_" + ext])
    run("make_boxes.py", *dims, "-pixel", dcell, "-nHDU", ns, "-ncpu", 2, "-PkDir", d["pk"], "-seed", int(g["seed"]),
        "-rsd", "True", "-outDir", d["boxes"], "-noise", "mt19937")
    assert rel_l2(np.load(d["boxes"] + "/boxk.npy"), g["boxkP0"]) < 1e-5
    for name, n in (("boxln_1", ns), ("boxln_2", ns), ("boxln_3", ns), ("vx", ns), ("vy", ns), ("vz", ns), ("box", NX),
                    ("eta_xx", NX), ("eta_yy", NX), ("eta_zz", NX), ("eta_xy", NX), ("eta_xz", NX), ("eta_yz", NX)):
        files = [d["boxes"] + "/%s-%d.fits" % (name, i) for i in range(n)]
        assert len(glob.glob(d["boxes"] + "/%s-*.fits" % name)) == n
        box = np.concatenate([fitsio.read(f) for f in files])
        assert rel_l2(box, g["box_" + name]) < 1e-5, name
        h = fitsio.read_header(files[0])
        assert (h["NX"], h["NY"], h["NZ"], h["DX"]) == (NX, NY, NZ, dcell)
        assert abs(h["sigma"] / float(g["sigma_" + name]) - 1) < 1e-4 and h["seed"] == int(g["seed"])
    write_qso_files(g, d["qso"])
    for i in range(ns):
        run("make_spectra.py", "-QSOfile", d["qso"] + "/QSO-", "-boxdir", d["boxes"], "-outDir", d["spectra"], "-i", i,
            "-N", ns, "-zmin", 1.8, "-zmax", 3.6, "-rsd", "True", "-dla", "True")
    ref_files = sorted(k[:-len("_THING_ID")] for k in g if k.startswith("spectra_") and k.endswith("_THING_ID"))
    got_files = sorted(os.path.basename(f).split(".")[0].replace("-", "_") for f in glob.glob(d["spectra"] + "/*.gz"))
    assert got_files == ref_files
    for key in ref_files:
        f = fitsio.FITS(d["spectra"] + "/" + key.replace("_", "-") + ".fits.gz")
        assert np.array_equal(f["METADATA"].read()["THING_ID"], g[key + "_THING_ID"])
        assert f["METADATA"].read_header()["Npixel"] == int(g[key + "_Npixel"])
        assert np.array_equal(f["LAMBDA"].read(), g[key + "_LAMBDA"])
        assert np.array_equal(f["REDSHIFT"].read(), g[key + "_REDSHIFT"])
        for ext, tol in (("DELTA_L", 1e-5), ("ETA_PAR", 1e-5), ("VELO_PAR", 2e-2)):
            assert np.max(np.abs(f[ext].read() - g[key + "_" + ext])) < tol, (key, ext)
    p1dfile = p1dmiss.build_pkmiss_interp(str(tmp_path / "pkmiss_standin.fits"))
    for i in range(ns):
        run("merge_spectra.py", "-inDir", d["spectra"], "-outDir", d["merged"], "-i", i, "-p1dfile", p1dfile, "-seed",
            int(g["seed"]), "-rsd", "True", "-dla", "True", "--store-g", "True", "-ncpu", 1, "-bb", -1, "-zfix", 2.4)
    ref = {k[len("merged_zfix_pixfile_"):]: v for k, v in g.items() if k.startswith("merged_zfix_pixfile_")}
    got = sorted(os.path.basename(f).split(".")[0] for f in glob.glob(d["merged"] + "/*.gz"))
    assert got == sorted(ref)
    ids = list(g["merged_zfix_THING_ID"])
    for name, tid in ref.items():
        f = fitsio.FITS(d["merged"] + "/" + name + ".fits.gz")
        assert np.array_equal(f["METADATA"].read()["THING_ID"], tid)
        rows = [ids.index(t) for t in tid]
        assert np.array_equal(f["LAMBDA"].read(), g["merged_zfix_LAMBDA"])
        assert np.max(np.abs(f["FLUX"].read() - g["merged_zfix_FLUX"][rows])) < 1e-5
        assert np.max(np.abs(f["DELTA_S"].read() - g["merged_zfix_DELTA_S"][rows])) < 2e-5
        assert np.max(np.abs(f["DELTA_L"].read() - g["merged_zfix_DELTA_L"][rows])) < 1e-5


def test_draw_qso_cli_matches_reference_files(tmp_path):
    """bin/draw_qso.py with the reference's CLI on the boxes the reference's make_boxes.py wrote: same quasars as the
    unmodified reference draw_qso.py (tests/golden/ref_qso.npz), same table layout."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from saclaymocks_b200 import fitsio_lite as fitsio
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_qso.npz")))
    NX, NY, NZ, dcell, ns = int(g["NX"]), int(g["NY"]), int(g["NZ"]), float(g["dcell"]), int(g["nslice"])
    boxes, out = str(tmp_path / "boxes"), str(tmp_path / "qso")
    os.makedirs(boxes), os.makedirs(out)
    for name in ("boxln_1", "boxln_2", "boxln_3", "vx", "vy", "vz"):
        for i in range(ns):
            f = fitsio.FITS(boxes + "/%s-%d.fits" % (name, i), "rw", clobber=True)
            f.write(g["box_" + name][i * NX // ns:(i + 1) * NX // ns],
                    header={"DX": dcell, "DY": dcell, "DZ": dcell, "NX": NX, "NY": NY, "NZ": NZ})
            f.close()
    for i in range(ns):
        log = run("draw_qso.py", "-indir", boxes, "-outpath", out, "-i", i, "-Nslice", ns, "-chunk", int(g["chunk"]), "-ra0",
                  float(g["ra0"]), "-dec0", float(g["dec0"]), "-dra", float(g["dra"]), "-ddec", float(g["ddec"]), "-zmin",
                  1.8, "-zmax", 3.6, "-desi", "False", "-seed", int(g["seed"]), "-rsd", "True")
        f = fitsio.FITS(out + "/QSO-%d-%d.fits" % (i, ns))
        t, hd = f[1].read(), f[1].read_header()
        assert hd["seed"] == int(g["seed"]) + i and hd["ra0"] == float(g["ra0"]) and hd["dec0"] == float(g["dec0"])
        assert "%d QSOs drawn" % len(g["qso%d_RA" % i]) in log
        for c in ("HDU", "THING_ID", "PLATE", "MJD", "FIBERID", "PMF", "XX"):
            assert np.array_equal(t[c], g["qso%d_%s" % (i, c)]), (i, c)
        for c in ("Z_QSO_NO_RSD", "Z_QSO_RSD", "RA", "DEC", "YY", "ZZ"):
            assert t[c].dtype == g["qso%d_%s" % (i, c)].dtype
            assert np.allclose(t[c], g["qso%d_%s" % (i, c)], rtol=3e-7, atol=0), (i, c)
    assert run.__name__ == "run"
    r = subprocess.run([sys.executable, os.path.join(BIN, "draw_qso.py"), "-indir", boxes, "-outpath", out, "-i", "0",
                        "-Nslice", str(ns), "-desi", "True"], capture_output=True, text=True)
    assert r.returncode != 0 and "not supported" in r.stdout          # fails loudly instead of silently skipping


def test_make_boxes_and_make_spectra_under_torchrun(tmp_path, golden_small):
    """The reference's CLIs on two GPUs: `torchrun bin/make_boxes.py` (box sharded over the ranks, every rank writes the
    files of its planes, weight tables evaluated on the GPU) and `torchrun bin/make_spectra.py` without -i (slices dealt
    to the ranks) write the same files as the reference run (tests/golden/ref_small.npz)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from saclaymocks_b200 import fitsio_lite as fitsio
    g = golden_small
    NX, NY, NZ, dcell, ns = int(g["NX"]), int(g["NY"]), int(g["NZ"]), float(g["dcell"]), int(g["nslice"])
    d = {k: str(tmp_path / k) for k in ("boxes", "qso", "spectra")}
    for v in d.values():
        os.makedirs(v)

    def torchrun(script, *argv):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
               "127.0.0.1", "--master-port", "29533", os.path.join(BIN, script)] + [str(a) for a in argv]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        return r.stdout

    torchrun("make_boxes.py", "-NX", NX, "-NY", NY, "-NZ", NZ, "-pixel", dcell, "-nHDU", ns, "-PkDir", "gpu", "-seed",
             int(g["seed"]), "-rsd", "True", "-outDir", d["boxes"], "-noise", "mt19937")
    for name, n in (("boxln_1", ns), ("boxln_2", ns), ("boxln_3", ns), ("vx", ns), ("vy", ns), ("vz", ns), ("box", NX),
                    ("eta_xx", NX), ("eta_yy", NX), ("eta_zz", NX), ("eta_xy", NX), ("eta_xz", NX), ("eta_yz", NX)):
        files = [d["boxes"] + "/%s-%d.fits" % (name, i) for i in range(n)]
        assert len(glob.glob(d["boxes"] + "/%s-*.fits" % name)) == n
        box = np.concatenate([fitsio.read(f) for f in files])
        assert rel_l2(box, g["box_" + name]) < 1e-5, name
        h = fitsio.read_header(files[0])
        assert (h["NX"], h["NY"], h["NZ"], h["DX"]) == (NX, NY, NZ, dcell)
        assert abs(h["sigma"] / float(g["sigma_" + name]) - 1) < 1e-4 and h["seed"] == int(g["seed"])
    assert len(glob.glob(d["boxes"] + "/boxk-*of2.npy")) == 2
    write_qso_files(g, d["qso"])
    torchrun("make_spectra.py", "-QSOfile", d["qso"] + "/QSO-", "-boxdir", d["boxes"], "-outDir", d["spectra"], "-N", ns,
             "-zmin", 1.8, "-zmax", 3.6, "-rsd", "True", "-dla", "True")
    ref_files = sorted(k[:-len("_THING_ID")] for k in g if k.startswith("spectra_") and k.endswith("_THING_ID"))
    got_files = sorted(os.path.basename(f).split(".")[0].replace("-", "_") for f in glob.glob(d["spectra"] + "/*.gz"))
    assert got_files == ref_files
    for key in ref_files:
        f = fitsio.FITS(d["spectra"] + "/" + key.replace("_", "-") + ".fits.gz")
        assert np.array_equal(f["METADATA"].read()["THING_ID"], g[key + "_THING_ID"])
        for ext, tol in (("DELTA_L", 1e-5), ("ETA_PAR", 1e-5), ("VELO_PAR", 2e-2)):
            assert np.max(np.abs(f[ext].read() - g[key + "_" + ext])) < tol, (key, ext)
