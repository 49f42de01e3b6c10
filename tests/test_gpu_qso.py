"""GPU parity of smk_draw_qso through the C ABI: fed the reference's legacy NumPy draws,
the kernel selects exactly the reference's quasars (same cells, in the same order) with float32-identical columns up
to the last bit of libm's atan/asin/pow; with Philox draws the selection is statistically equivalent."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from test_qso_cpu import COLS, args, slice_boxes  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_qso.npz")


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def run_gpu(cuda, boxes, a, i, uniforms_seed=None, philox_seed=0):
    from saclaymocks_b200 import qso
    from saclaymocks_b200.boxes import BoxSynth
    NXs, NY, NZ = boxes["boxln_1"].shape
    bs = BoxSynth(16, 16, 24, 2.19, device=cuda)           # only its ctx / stream are used
    dev = {k: torch.as_tensor(np.ascontiguousarray(v, dtype=np.float32), device=cuda) for k, v in boxes.items()}
    sig = tuple(qso.box_sigma(dev[k]) for k in ("boxln_1", "boxln_2", "boxln_3"))
    sig = tuple(float(np.float32(s)) for s in sig)           # np.std of a float32 array is float32
    st = qso.QsoSetup(NXs, NY, NZ, a["NX_full"], a["dcell"], i, a["nslice"], a["ra0"], a["dec0"], a["dra"], a["ddec"],
                      a["zmin"], a["zmax"], sig)
    uni, rs = (None, None)
    if uniforms_seed is not None:
        uni, rs = qso.legacy_uniforms(uniforms_seed, NXs, NY, NZ)
    d = qso.QsoDrawer(bs)
    cat = d.draw(st, [dev["boxln_1"], dev["boxln_2"], dev["boxln_3"]], [dev["vx"], dev["vy"], dev["vz"]],
                 ix0=i * NXs, uniforms=uni, seed=philox_seed, chunk=a["chunk"], rs=rs)
    bs.close()
    return cat


def test_gpu_reproduces_the_reference_catalogue(cuda):
    g = dict(np.load(GOLDEN))
    for i in range(int(g["nslice"])):
        a = args(g, i)
        cat = run_gpu(cuda, slice_boxes(g, i), a, i, uniforms_seed=a["seed"] + i)
        for c in ("HDU", "THING_ID", "PLATE", "MJD", "FIBERID", "PMF"):
            assert np.array_equal(cat[c], g["qso%d_%s" % (i, c)]), (i, c)          # same count, same stream
        for c in ("Z_QSO_NO_RSD", "Z_QSO_RSD", "RA", "DEC", "XX", "YY", "ZZ"):
            ref = g["qso%d_%s" % (i, c)]
            assert np.allclose(cat[c], ref, rtol=3e-7, atol=0), (i, c, np.abs(cat[c] - ref).max())
        assert np.array_equal(cat["XX"], g["qso%d_XX" % i])                          # pure + and *: bit-exact


def test_gpu_selection_matches_oracle_on_a_larger_slab(cuda):
    """64 x 64 x 384 cells (1.6e6 cells, ~150 quasars): identical cell indices, in order; float64 records within 1e-12."""
    from oracle import draw_qso as odq
    rng = np.random.default_rng(5)
    NXs, NY, NZ, dcell = 32, 64, 384, 8.76
    boxes = {k: (0.9 * rng.standard_normal((NXs, NY, NZ))).astype(np.float32) for k in ("boxln_1", "boxln_2", "boxln_3")}
    boxes.update({k: (300 * rng.standard_normal((NXs, NY, NZ))).astype(np.float32) for k in ("vx", "vy", "vz")})
    a = dict(NX_full=64, dcell=dcell, i_slice=1, nslice=2, chunk=3, ra0=190.0, dec0=5.0, dra=30.0, ddec=30.0, zmin=1.8,
             zmax=3.6, seed=11)
    ref = odq.draw_qso_slice(boxes, **a)
    cat = run_gpu(cuda, boxes, a, 1, uniforms_seed=a["seed"] + 1)
    assert len(ref["RA"]) > 50
    assert cat["nn_cond1"] == ref["nn_cond1"]
    assert np.array_equal(cat["cells"], ref["cells"])
    for j, k in enumerate(("z", "zrsd", "ra", "dec", "xx", "yy", "zz")):
        assert np.allclose(cat["f64"][:, j + 1], ref["f64"][k], rtol=1e-12, atol=1e-12), k
    for c in COLS:
        if c in ("HDU", "THING_ID", "PLATE", "MJD", "FIBERID", "PMF"):
            assert np.array_equal(cat[c], ref[c]), c


def test_gpu_philox_selection_rate(cuda):
    """Philox draws: same expected number of quasars as the NumPy stream (Poisson error), reproducible, and
    independent of the uniform arrays' absence (no host stream needed)."""
    rng = np.random.default_rng(6)
    NXs, NY, NZ, dcell = 32, 64, 384, 8.76
    boxes = {k: (0.9 * rng.standard_normal((NXs, NY, NZ))).astype(np.float32) for k in ("boxln_1", "boxln_2", "boxln_3")}
    boxes.update({k: (300 * rng.standard_normal((NXs, NY, NZ))).astype(np.float32) for k in ("vx", "vy", "vz")})
    a = dict(NX_full=64, dcell=dcell, i_slice=0, nslice=2, chunk=3, ra0=190.0, dec0=5.0, dra=30.0, ddec=30.0, zmin=1.8,
             zmax=3.6, seed=11)
    n_np = len(run_gpu(cuda, boxes, a, 0, uniforms_seed=5)["RA"])
    c1 = run_gpu(cuda, boxes, a, 0, philox_seed=77)
    c2 = run_gpu(cuda, boxes, a, 0, philox_seed=77)
    c3 = run_gpu(cuda, boxes, a, 0, philox_seed=78)
    assert np.array_equal(c1["cells"], c2["cells"]) and np.array_equal(c1["f64"], c2["f64"])
    assert not np.array_equal(c1["cells"], c3["cells"])
    n1, n3 = len(c1["RA"]), len(c3["RA"])
    assert abs(n1 - n_np) < 6 * np.sqrt(n_np) + 5 and abs(n3 - n_np) < 6 * np.sqrt(n_np) + 5


def test_gpu_empty_selection_and_record_overflow(cuda):
    """No quasar in the redshift window -> empty, well-formed table; a record buffer that is too small is detected
    through counters[0] and the call repeated (same catalogue as with a large buffer)."""
    from saclaymocks_b200 import qso
    from saclaymocks_b200.boxes import BoxSynth
    rng = np.random.default_rng(8)
    NXs, NY, NZ, dcell = 16, 32, 384, 8.76
    boxes = {k: (0.9 * rng.standard_normal((NXs, NY, NZ))).astype(np.float32) for k in ("boxln_1", "boxln_2", "boxln_3")}
    boxes.update({k: (300 * rng.standard_normal((NXs, NY, NZ))).astype(np.float32) for k in ("vx", "vy", "vz")})
    a = dict(NX_full=32, dcell=dcell, i_slice=0, nslice=2, chunk=1, ra0=190.0, dec0=5.0, dra=1e-7, ddec=1e-7, zmin=1.8,
             zmax=3.6, seed=3)                                   # an angular window no cell centre falls into
    cat = run_gpu(cuda, boxes, a, 0, philox_seed=1)
    assert len(cat["RA"]) == 0 and cat["PMF"].dtype == np.dtype("S21") and cat["cells"].shape == (0, 3)
    assert cat["nn_cond1"] > 0                                   # cells did pass cond1; cond3 removed them all
    a["dra"] = a["ddec"] = 30.0
    bs = BoxSynth(16, 16, 24, 2.19, device=cuda)
    dev = {k: torch.as_tensor(v, device=cuda) for k, v in boxes.items()}
    sig = tuple(float(np.float32(qso.box_sigma(dev[k]))) for k in ("boxln_1", "boxln_2", "boxln_3"))
    st = qso.QsoSetup(NXs, NY, NZ, a["NX_full"], dcell, 0, 2, a["ra0"], a["dec0"], a["dra"], a["ddec"], 1.8, 3.6, sig,
                      rho_sum=qso.scaled_rho_sum(NXs * NY * NZ) / 50)          # ~50x the nominal density
    d = qso.QsoDrawer(bs)
    args_ = ([dev["boxln_1"], dev["boxln_2"], dev["boxln_3"]], [dev["vx"], dev["vy"], dev["vz"]])
    big = d.draw(st, *args_, seed=5)
    small = d.draw(st, *args_, seed=5, capacity=3)
    assert len(big["RA"]) > 10 and np.array_equal(big["f64"], small["f64"])
    bs.close()


def test_gpu_fast_kernel_probability_matches_oracle_ptot(cuda):
    """Production (Philox) kernel: its float32 / tabulated ptot gives the cond1 acceptance the oracle's float64 ptot
    predicts -- sum over cells of min(norm * ptot, 1) -- within the binomial error, at ~50x the nominal density."""
    from oracle import draw_qso as odq
    from saclaymocks_b200 import qso
    from saclaymocks_b200.boxes import BoxSynth
    rng = np.random.default_rng(9)
    NXs, NY, NZ, dcell = 32, 64, 384, 8.76
    boxes = {k: (0.9 * rng.standard_normal((NXs, NY, NZ))).astype(np.float32) for k in ("boxln_1", "boxln_2", "boxln_3")}
    boxes.update({k: (300 * rng.standard_normal((NXs, NY, NZ))).astype(np.float32) for k in ("vx", "vy", "vz")})
    sig = tuple(float(np.std(boxes[k])) for k in ("boxln_1", "boxln_2", "boxln_3"))
    common = (NXs, NY, NZ, 64, dcell, 1, 2, 190.0, 5.0, 30.0, 30.0, 1.8, 3.6, sig)
    st = qso.QsoSetup(*common, rho_sum=qso.scaled_rho_sum(NXs * NY * NZ) / 50)
    ost = odq.Setup(*common)
    # P(cond1) = clip(norm * ptot, 0, 1): the linear z interpolation extrapolates to negative weights outside
    # [z_QSO_bias_1, z_QSO_bias_3], so ptot can be negative (never selected)
    expected = float(np.clip(st.norm * odq.ptot_box(ost, boxes["boxln_1"], boxes["boxln_2"], boxes["boxln_3"]), 0, 1).sum())
    bs = BoxSynth(16, 16, 24, 2.19, device=cuda)
    dev = {k: torch.as_tensor(v, device=cuda) for k, v in boxes.items()}
    d = qso.QsoDrawer(bs)
    nn = [d.draw(st, [dev["boxln_1"], dev["boxln_2"], dev["boxln_3"]], [dev["vx"], dev["vy"], dev["vz"]], ix0=NXs,
                 seed=s)["nn_cond1"] for s in (1, 2, 3, 4)]
    assert expected > 2e4
    assert abs(np.mean(nn) - expected) < 5 * np.sqrt(expected / len(nn)), (nn, expected)
    assert len(set(nn)) > 1                                                  # different seeds, different draws
    bs.close()


def test_gpu_fast_kernel_selects_the_exact_kernels_quasars(cuda):
    """Production kernel (tabulated float32 ptot) against the reference-arithmetic kernel on the SAME Philox variates
    (both take cond1's variate from cond1_block(), smk_qso.cu): the two selections are identical except for cells
    whose variate falls inside the float32 / table error of norm * ptot (relative 1e-4: a handful out of ~1e4 cond1
    survivors); every quasar selected by both has bit-identical float64 records (same finish_cell())."""
    from saclaymocks_b200 import qso
    from saclaymocks_b200.boxes import BoxSynth
    rng = np.random.default_rng(10)
    NXs, NY, NZ, dcell = 32, 64, 384, 8.76
    boxes = {k: (0.9 * rng.standard_normal((NXs, NY, NZ))).astype(np.float32) for k in ("boxln_1", "boxln_2", "boxln_3")}
    boxes.update({k: (300 * rng.standard_normal((NXs, NY, NZ))).astype(np.float32) for k in ("vx", "vy", "vz")})
    sig = tuple(float(np.std(boxes[k])) for k in ("boxln_1", "boxln_2", "boxln_3"))
    st = qso.QsoSetup(NXs, NY, NZ, 64, dcell, 1, 2, 190.0, 5.0, 30.0, 30.0, 1.8, 3.6, sig,
                      rho_sum=qso.scaled_rho_sum(NXs * NY * NZ) / 50)          # ~50x the nominal density
    bs = BoxSynth(16, 16, 24, 2.19, device=cuda)
    dev = {k: torch.as_tensor(v, device=cuda) for k, v in boxes.items()}
    d = qso.QsoDrawer(bs)
    a = ([dev["boxln_1"], dev["boxln_2"], dev["boxln_3"]], [dev["vx"], dev["vy"], dev["vz"]])
    fast = d.draw(st, *a, ix0=NXs, seed=21)
    from saclaymocks_b200 import _lib
    with _lib.option("qso_exact", 1):
        exact = d.draw(st, *a, ix0=NXs, seed=21)
    kf = {tuple(c): i for i, c in enumerate(fast["cells"])}
    ke = {tuple(c): i for i, c in enumerate(exact["cells"])}
    common = sorted(set(kf) & set(ke))
    odd = set(kf) ^ set(ke)
    assert len(exact["RA"]) > 3000
    assert len(odd) <= max(3, 2e-3 * len(ke)), (len(odd), len(ke))
    assert abs(fast["nn_cond1"] - exact["nn_cond1"]) <= max(5, 2e-3 * exact["nn_cond1"])
    fi, ei = [kf[c] for c in common], [ke[c] for c in common]
    assert np.array_equal(fast["f64"][fi], exact["f64"][ei])
    bs.close()
