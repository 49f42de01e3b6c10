"""GPU parity and size-independent properties at the box sizes of BASELINE.json's configs: the x / y transform
lengths 256, 512, 1024, 2048, 2560 (radix plans 16.16, 8.8.8, 16.16.4, 16.16.8, 16.16.2.5) against the CPU oracle on
thin boxes the oracle finishes in seconds, and the bench workload's own size (512 x 512 x 1536, config 2) through
properties that need no oracle: round trip, Parseval, and the Gaussian-weighted average of constant fields."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import rel_l2  # noqa: E402

TOL = 1e-5          # relative L2 on every box (BASELINE.json north_star)


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("shape", [(256, 512, 24), (512, 16, 96), (1024, 16, 24), (16, 1024, 24), (2048, 16, 24),
                                   (16, 2048, 24), (2560, 16, 24), (16, 2560, 24)])
def test_long_axis_plans_match_oracle(cuda, shape):
    """Every product of a thin box whose x or y length is one of the configs' (all fused-multiply variants of the x
    pass, the plain y pass, both directions) against the oracle, fed the reference's MT19937 noise."""
    from oracle import boxes as ob
    from oracle import pk_weights
    from saclaymocks_b200.boxes import BoxSynth, PRODUCTS, WEIGHT_OF
    NX, NY, NZ = shape
    dcell = 2.19
    W = pk_weights.weights(NX, NY, NZ, dcell)
    noise = ob.draw_noise(NX, NY, NZ, 42)
    raw, p0, boxes, sig = ob.make_boxes(NX, NY, NZ, dcell, 42, W, workers=8, noise=noise)
    bs = BoxSynth(NX, NY, NZ, dcell, device=cuda)
    boxk = bs.draw_grf_boxk(noise=torch.as_tensor(noise, device=cuda))
    assert rel_l2(bs.boxk_to_numpy(boxk), raw) < TOL
    Wd = {k: bs.upload_weights(v) for k, v in W.items()}
    for name in PRODUCTS:
        box, stats = bs.synth(boxk, name, wtable=Wd.get(WEIGHT_OF.get(name)))
        assert rel_l2(box.cpu().numpy(), boxes[name]) < TOL, name
        assert abs(bs.sigma(stats) / sig[name] - 1) < 1e-4, name
        if name == "box":
            assert rel_l2(bs.boxk_to_numpy(boxk), p0) < TOL
    bs.close()


def test_roundtrip_and_parseval_at_bench_size(cuda):
    """512 x 512 x 1536 (BASELINE config 2, the bench workload): c2r(r2c(x)) / N == x with unit weights, Parseval with
    Hermitian multiplicities, and the sums the c2r pass accumulates for sigma."""
    from saclaymocks_b200.boxes import BoxSynth
    NX, NY, NZ = 512, 512, 1536
    bs = BoxSynth(NX, NY, NZ, 2.19, device=cuda)
    x = bs.noise_philox(11)
    boxk = bs.draw_grf_boxk(noise=x)
    ones = torch.ones((NX, NY, NZ // 2 + 1), dtype=torch.float32, device=cuda)
    y, stats = bs.synth(boxk, "boxln_1", wtable=ones)
    num = torch.zeros((), dtype=torch.float64, device=cuda)
    den = torch.zeros((), dtype=torch.float64, device=cuda)
    s1 = torch.zeros((), dtype=torch.float64, device=cuda)
    for i in range(0, NX, 64):                     # chunked: no 3 GB float64 temporaries
        xd, yd = x[i:i + 64].double(), y[i:i + 64].double()
        num += ((yd - xd) ** 2).sum()
        den += (xd ** 2).sum()
        s1 += yd.sum()
    assert float(torch.sqrt(num / den)) < 2e-6
    tot = torch.zeros((), dtype=torch.float64, device=cuda)
    for i in range(0, NX, 64):
        k2 = boxk[i:i + 64, :, :NZ // 2 + 1].abs().double() ** 2
        tot += 2 * k2.sum() - k2[:, :, 0].sum() - k2[:, :, NZ // 2].sum()
    assert abs(float(tot / (den * NX * NY * NZ)) - 1) < 1e-5
    st = stats.cpu().numpy()
    assert abs(st[0] - float(s1)) < 1e-3 * np.sqrt(NX * NY * NZ)        # sum (float32 partial sums per thread)
    assert abs(st[1] / float(den) - 1) < 1e-5                           # sum of squares
    bs.close()
    del x, y, boxk, ones
    torch.cuda.empty_cache()


def test_skewers_of_constant_fields_at_bench_size(cuda):
    """The bench workload's skewer stage (512 x 512 x 1536 box, its 14.7k full-density sightlines, 10 fields) on
    constant fields: a weighted average of a constant is the constant, so delta_l = c0, eta_par = c1 (eta_ij = c1
    delta_ij contracts with x_i x_j / r^2 to c1) and v_par = cz Z/R (v = cz e_z) at every computed pixel, whatever the
    window (interior fast path or clamped at the faces); pixels past the forest get the reference's sentinels."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from bench import synthetic_qsos
    from saclaymocks_b200 import spectra as sp
    NX = NY = 512
    NZ, dcell = 1536, 2.19
    geom = sp.SkewerGeometry(NX, NY, NZ, dcell)
    eng = sp.SkewerEngine(geom, device=cuda)
    ra, dec, z, ra0, dec0 = synthetic_qsos(NX, NY)
    xyzr, nfor = sp.qso_lines_of_sight(geom, ra, dec, z, ra0, dec0)
    keep = nfor >= 0
    xyzr, nfor = xyzr[keep], nfor[keep]
    c0, c1, cz = 0.75, -0.4, 120.0
    full = lambda v: torch.full((NX, NY, NZ), v, dtype=torch.float32, device=cuda)       # noqa: E731
    A, B, Z0, V = full(c0), full(c1), full(0.0), full(cz)
    fields = {"box": A, "eta_xx": B, "eta_yy": B, "eta_zz": B, "eta_xy": Z0, "eta_xz": Z0, "eta_yz": Z0, "vx": Z0,
              "vy": Z0, "vz": V}
    dl, ep, vp = eng.read_spec(fields, xyzr, nfor)
    npix = geom.npixeltot
    bad = 0
    ncomp = 0
    for i0 in range(0, len(nfor), 2048):
        sl = slice(i0, i0 + 2048)
        X = xyzr[sl, 0:1] * geom.R_vec[None, :] / xyzr[sl, 3:4]
        owned = (X > -geom.LX / 2) & (X <= geom.LX / 2)
        forest = np.arange(npix)[None, :] < nfor[sl, None]
        comp, past = owned & forest, owned & ~forest
        d, e, v = (t[sl].cpu().numpy() for t in (dl, ep, vp))
        zr = np.broadcast_to((xyzr[sl, 2] / xyzr[sl, 3])[:, None], d.shape)
        bad += int((np.abs(d[comp] - c0) > 1e-5 * abs(c0)).sum())
        bad += int((np.abs(e[comp] - c1) > 1e-5 * abs(c1)).sum())
        bad += int((np.abs(v[comp] - cz * zr[comp]) > 1e-5 * cz).sum())
        bad += int((d[past] != -1e6).sum() + (e[past] != 0).sum() + (v[past] != 0).sum())
        bad += int((~np.isnan(d[~owned])).sum())                      # pixels of no slab are left untouched
        ncomp += int(comp.sum())
    assert ncomp > 2e7 and bad == 0, (ncomp, bad)
    del A, B, Z0, V, dl, ep, vp
    torch.cuda.empty_cache()


def test_run_chunk_resident_chain(cuda):
    """ChunkPipeline.run_chunk on the reference's debugging box (chunk_parameters(32): 32 x 32 x 1536 cells, window
    190 +- 1.6 deg): boxes -> quasars drawn on the resident boxes -> sightlines -> FGPA without leaving the GPU.  The
    catalogue obeys the chunk window and draw_qso.py's THING_ID rule, and the rows equal the oracle's ReadSpec run on
    the same (downloaded) boxes for the same quasars."""
    from oracle import spectra as osp
    from saclaymocks_b200 import chunks
    from saclaymocks_b200.chunk import ChunkPipeline
    NX = NY = 32
    NZ, dcell = 1536, 2.19
    pipe = ChunkPipeline(NX, NY, NZ, dcell, device=cuda, zfix=2.4)
    pipe.set_weights({k: pipe.bs.weight_table(k) for k in ("Pln1", "Pln2", "Pln3", "P0")})
    cat, out = pipe.run_chunk(chunk=1, seed=5)
    ra0, dra, dec0, ddec = chunks.chunk_window(NX, 1)
    n = len(cat["RA"])
    assert n > 20, n
    assert np.array_equal(cat["THING_ID"], 10 ** 9 + np.arange(n) + 1)                  # draw_qso.py:495, slab 0
    assert np.all(np.abs(cat["RA"] - ra0) < dra) and np.all(np.abs(cat["DEC"] - dec0) < ddec)
    assert np.all((cat["Z_QSO_RSD"] > 1.8) & (cat["Z_QSO_RSD"] < 3.6))
    dl, ep, vp, F = (t.cpu().numpy() for t in out)
    assert dl.shape == (len(pipe.cat["sel"]), pipe.geom.npixeltot)
    boxes = {k: pipe.interior(k).cpu().numpy() for k in osp.FIELDS}
    og = osp.Geometry(NX, NY, NZ, dcell)
    half = np.degrees(np.arctan((og.LX / 2 - 3 * dcell) / (og.R0 + og.LZ / 2))) * 0.95     # 7^3 window inside the box
    safe = np.where((np.abs(cat["RA"] - ra0) < half) & (np.abs(cat["DEC"] - dec0) < half))[0][:25]
    assert len(safe) >= 3, len(safe)
    q = np.zeros(len(safe), dtype=[("RA", "f4"), ("DEC", "f4"), ("Z_QSO_NO_RSD", "f4"), ("Z_QSO_RSD", "f4"),
                                   ("THING_ID", "i8"), ("HDU", "i4")])
    for c in ("RA", "DEC", "Z_QSO_NO_RSD", "Z_QSO_RSD", "THING_ID"):
        q[c] = cat[c][safe]
    ids = list(pipe.cat["ids"])
    lam32 = np.float32(pipe.geom.lambda_vec)
    pieces = osp.make_spectra_slice(og, boxes, [q], 0, 1, ra0, dec0)
    assert len(pieces) >= 3
    checked = 0
    for p in pieces:
        row = ids.index(p["id"])
        idx = np.searchsorted(lam32, p["lam"])
        m = p["delta_l"] > -1e5
        assert np.all(dl[row, idx][~m] == -1e6)
        assert np.max(np.abs(ep[row, idx] - p["eta_par"])) < 1e-5
        if m.any():
            assert np.max(np.abs(dl[row, idx][m] - p["delta_l"][m])) < 1e-5
            f = F[row, idx][m]
            assert np.all((f >= 0) & (f <= 1))           # float32 F underflows to 0 in dense absorbers
            checked += int(m.sum())
    assert checked > 1000
    # the FGPA fused into the gather's epilogue (smk_skewers_fgpa, what step_skewers runs) against the separate pass
    # (smk_fgpa) over the same rows; pixels of no slab stay untouched in both (NaN after the refill)
    for t in pipe.out:
        t.fill_(float("nan"))
    dl_t, ep_t, vp_t, F_t = pipe.step_skewers(seed=5)
    F_sep = pipe.fgpa.flux(dl_t, pipe.delta_s, ep_t)
    assert torch.isfinite(F_t).sum() > 1000 and bool((torch.isfinite(F_t) == torch.isfinite(dl_t)).all())
    assert float((torch.nan_to_num(F_t, nan=2.0) - torch.nan_to_num(F_sep, nan=2.0)).abs().max()) < 1e-6
