"""draw_qso on the CPU: the oracle restatement is pinned against the unmodified
reference script (tests/golden/ref_qso.npz, made by tests/golden/run_reference_shimmed.py qso), and the product's host
set-up (saclaymocks_b200/qso.py) agrees with the oracle's to the last bit."""
import os

import numpy as np
import pytest

from oracle import draw_qso as odq

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_qso.npz")
COLS = ("Z_QSO_NO_RSD", "Z_QSO_RSD", "RA", "DEC", "HDU", "THING_ID", "PLATE", "MJD", "FIBERID", "PMF", "XX", "YY", "ZZ")


@pytest.fixture(scope="module")
def golden_qso():
    return dict(np.load(GOLDEN))


def slice_boxes(g, i):
    nxs = int(g["NX"]) // int(g["nslice"])
    return {k: g["box_" + k][i * nxs:(i + 1) * nxs] for k in ("boxln_1", "boxln_2", "boxln_3", "vx", "vy", "vz")}


def args(g, i):
    return dict(NX_full=int(g["NX"]), dcell=float(g["dcell"]), i_slice=i, nslice=int(g["nslice"]), chunk=int(g["chunk"]),
                ra0=float(g["ra0"]), dec0=float(g["dec0"]), dra=float(g["dra"]), ddec=float(g["ddec"]),
                zmin=float(g["zmin"]), zmax=float(g["zmax"]), seed=int(g["seed"]))


def test_oracle_reproduces_the_reference_catalogue(golden_qso):
    """Every column of QSO-<i>-<N>.fits, bit for bit (selection, sub-cell positions, RSD redshift, MJD/FIBERID)."""
    g = golden_qso
    total = 0
    for i in range(int(g["nslice"])):
        r = odq.draw_qso_slice(slice_boxes(g, i), **args(g, i))
        assert int(g["qso%d_seed" % i]) == int(g["seed"]) + i          # draw_qso.py:154
        for c in COLS:
            ref = g["qso%d_%s" % (i, c)]
            assert r[c].shape == ref.shape and np.all(r[c] == ref), (i, c)
        total += len(r["RA"])
    assert total >= 20


def test_host_setup_matches_oracle(golden_qso):
    from saclaymocks_b200 import qso
    g = golden_qso
    a = args(g, 1)
    b = slice_boxes(g, 1)
    sig = tuple(np.std(b[k]) for k in ("boxln_1", "boxln_2", "boxln_3"))
    NXs, NY, NZ = b["boxln_1"].shape
    o = odq.Setup(NXs, NY, NZ, a["NX_full"], a["dcell"], 1, a["nslice"], a["ra0"], a["dec0"], a["dra"], a["ddec"],
                  a["zmin"], a["zmax"], sig)
    s = qso.QsoSetup(NXs, NY, NZ, a["NX_full"], a["dcell"], 1, a["nslice"], a["ra0"], a["dec0"], a["dra"], a["ddec"],
                     a["zmin"], a["zmax"], sig)
    assert s.norm == o.norm and s.density_max == o.density_max
    assert (s.z_min, s.z_max, s.dgrowth0) == (o.z_min, o.z_max, o.dgrowth0)
    for k in ("x_axis", "y_axis", "z_axis", "dn_cell", "dz_interp", "coef_z", "coef_v"):
        assert np.array_equal(getattr(s, k), getattr(o, k)), k
    (u, _), (v, _) = qso.legacy_uniforms(43, 4, 5, 6), odq.draw_uniforms(43, 4, 5, 6)
    assert all(np.array_equal(x, y) for x, y in zip(u, v))
    zz = np.linspace(1.81, 3.59, 50)
    assert np.array_equal(s.cond1_correction(zz), o.cond1_correction(zz))


def test_known_answers_qso():
    """Anchors computed from the reference formulas: a(z_b, z_b) = 1; the coefficient table is 1 at z <= 1.9 and 0
    above its last entry; diffmod wraps like Python's %."""
    assert odq.qso_a_of_z(2.75, 2.75) == 1.0
    z, c = odq.lognormal_coef_table()
    assert c[0] == 1.0 and c[1] == 1.0 and c[-1] == 0.0 and z[1] == 1.9 and z[-1] == 10.0
    assert odq.diffmod(359.0, 1.0, 360) == 2.0 and odq.diffmod(-1.0, 1.0, 180) == 2.0
