"""GPU parity of the skewer stage (gather, small-scale field, FGPA) against the oracle and the reference goldens.
Tolerances (BASELINE.json north_star): transmissions within 1e-5 absolute."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import qso_files_from_golden  # noqa: E402


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _gpu_skewers(g, boxes, cuda, rsd=True, dla=True, slabs=1):
    from saclaymocks_b200 import spectra as sp
    NX, NY, NZ, dcell = int(g["NX"]), int(g["NY"]), int(g["NZ"]), float(g["dcell"])
    geom = sp.SkewerGeometry(NX, NY, NZ, dcell)
    q = np.concatenate(qso_files_from_golden(g))
    zq = q["Z_QSO_RSD"] if rsd else q["Z_QSO_NO_RSD"]
    xyzr, nfor = sp.qso_lines_of_sight(geom, q["RA"], q["DEC"], zq, float(g["ra0"]), float(g["dec0"]))
    eng = sp.SkewerEngine(geom, device=cuda)
    out = None
    for s in range(slabs):          # x-slab decomposition with dmax halo planes, like make_spectra -i s -N slabs
        lo, hi = max(s * NX // slabs - geom.dmax, 0), min((s + 1) * NX // slabs + geom.dmax, NX)
        f = {k: torch.as_tensor(np.ascontiguousarray(boxes[k][lo:hi]), device=cuda) for k in sp.FIELDS}
        xmin, xmax = geom.LX * s / slabs - geom.LX / 2, geom.LX * (s + 1) / slabs - geom.LX / 2
        out = eng.read_spec(f, xyzr, np.maximum(nfor, 0), ix0=lo, xmin=xmin, xmax=xmax, rsd=rsd, dla=dla, out=out)
    return geom, q, xyzr, nfor, out


def _compare_with_pieces(g, geom, q, out):
    """Golden make_spectra pieces (per slab) against the full-row GPU result, matched by wavelength."""
    d, e, v = (t.cpu().numpy() for t in out)
    v = v * np.float32(1) * geom.velo_rescale()[None, :]            # make_spectra.py:510
    lam32 = np.float32(geom.lambda_vec)
    ids = list(q["THING_ID"])
    n = 0
    for key in [k[:-len("_THING_ID")] for k in g if k.startswith("spectra_") and k.endswith("_THING_ID")]:
        for r, ID in enumerate(g[key + "_THING_ID"]):
            lam = g[key + "_LAMBDA"][r]
            m = lam > 0
            idx = np.searchsorted(lam32, lam[m])
            assert np.array_equal(lam32[idx], lam[m])
            row = ids.index(ID)
            for ext, arr, tol in (("DELTA_L", d, 5e-6), ("ETA_PAR", e, 5e-6), ("VELO_PAR", v, 2e-3)):
                ref = g[key + "_" + ext][r][m]
                assert np.max(np.abs(arr[row, idx] - ref)) < tol * max(1.0, np.abs(ref[ref > -1e5]).max()
                                                                      if np.any(ref > -1e5) else 1.0), (key, ID, ext)
            n += 1
    assert n > 0


def test_skewers_match_reference_small(cuda, golden_small):
    g = golden_small
    from saclaymocks_b200 import spectra as sp
    boxes = {k: g["box_" + k] for k in sp.FIELDS}
    geom, q, xyzr, nfor, out = _gpu_skewers(g, boxes, cuda)
    _compare_with_pieces(g, geom, q, out)


def test_skewers_slab_decomposition_is_exact(cuda, golden_small):
    """Sharding by the slab that owns the pixel (make_spectra.py:443-448) gives bit-identical rows."""
    g = golden_small
    from saclaymocks_b200 import spectra as sp
    boxes = {k: g["box_" + k] for k in sp.FIELDS}
    _, _, _, _, full = _gpu_skewers(g, boxes, cuda, slabs=1)
    _, _, _, _, slab = _gpu_skewers(g, boxes, cuda, slabs=4)
    for a, b in zip(full, slab):
        assert torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))


def test_skewers_vs_oracle_no_dla_no_rsd(cuda, golden_small):
    from oracle import spectra as osp
    from saclaymocks_b200 import spectra as sp
    g = golden_small
    boxes = {k: g["box_" + k] for k in sp.FIELDS}
    for rsd, dla in ((True, False), (False, False)):
        geom, q, xyzr, nfor, out = _gpu_skewers(g, boxes, cuda, rsd=rsd, dla=dla)
        og = osp.Geometry(int(g["NX"]), int(g["NY"]), int(g["NZ"]), float(g["dcell"]))
        qf = qso_files_from_golden(g)
        ns = int(g["nslice"])
        d = out[0].cpu().numpy()
        e = out[1].cpu().numpy()
        ids = list(q["THING_ID"])
        lam32 = np.float32(geom.lambda_vec)
        for s in range(ns):
            for p in osp.make_spectra_slice(og, boxes, qf, s, ns, float(g["ra0"]), float(g["dec0"]), rsd=rsd, dla=dla):
                idx = np.searchsorted(lam32, p["lam"])
                row = ids.index(p["id"])
                assert np.max(np.abs(d[row, idx] - p["delta_l"])) < 5e-6 * 1e0 or np.all(p["delta_l"] < -1e5)
                assert np.max(np.abs(e[row, idx] - p["eta_par"])) < 5e-6


def test_end_to_end_ref32(cuda, golden_ref32):
    """32 x 32 x 1536 reference run: MT19937 noise -> GPU boxes -> GPU skewers -> GPU delta_s (reference noise
    stream) + FGPA -> FLUX of the unmodified merge_spectra.py within 1e-5 absolute."""
    _end_to_end(cuda, golden_ref32)


def test_end_to_end_c1(cuda, golden_c1):
    """The same chain at BASELINE config 1's full size (256 x 256 x 1536 cells, 8 slices): GPU boxes against the
    reference's FITS boxes on a strided sample (1e-5 relative L2), GPU skewers against its spectra pieces, GPU FLUX
    against the unmodified merge_spectra.py within 1e-5 absolute."""
    _end_to_end(cuda, golden_c1)
    torch.cuda.empty_cache()


def test_end_to_end_c2(cuda, golden_c2):
    """The same chain at the bench workload's size (BASELINE config 2, 512 x 512 x 1536) against the unmodified
    reference's files: MT19937 noise from the host (14 s), weight table evaluated on the GPU and checked against the
    reference's P-file sample, boxes / skewer pieces / FLUX within the north-star tolerances.  Runs in the default set."""
    _end_to_end(cuda, golden_c2)
    torch.cuda.empty_cache()


def _end_to_end(cuda, g):
    from oracle import boxes as ob
    from oracle import pk_weights
    from oracle import merge as om
    from saclaymocks_b200 import spectra as sp
    from saclaymocks_b200.boxes import BoxSynth, WEIGHT_OF
    NX, NY, NZ, dcell = int(g["NX"]), int(g["NY"]), int(g["NZ"]), float(g["dcell"])
    # the reference's MT19937 stream, drawn plane-major on the host (contiguous writes) and permuted on the device
    planes = torch.as_tensor(ob.draw_noise_planes(NX, NY, NZ, int(g["seed"])), device=cuda)
    noise = planes.permute(1, 2, 0).contiguous()
    del planes
    bs = BoxSynth(NX, NY, NZ, dcell, device=cuda)
    boxk = bs.draw_grf_boxk(noise=noise)
    del noise
    if NX * NY * NZ > 2e8:      # config 2: the host spline evaluation of 8e8 table entries takes minutes; the GPU tables
        Wd = {"P0": bs.weight_table("P0")}                                       # are bit-equal to > 99.9 %, 1 ulp else
        st = int(g["stride"])
        for k, t in Wd.items():
            got = t.cpu().numpy().ravel()[::st]
            ref = g["W_" + k]
            assert (got == ref).mean() > 0.999, k
            assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-30)) < 2.5e-7, k     # W(k = 0) is exactly 0
    else:
        W = pk_weights.weights(NX, NY, NZ, dcell)
        Wd = {"P0": bs.upload_weights(W["P0"])}
    # reference order (make_boxes.py:289-431): 'box' stores boxk*P0 back, the eta / velocity products read that
    fields = {"box": bs.synth(boxk, "box", wtable=Wd["P0"])[0]}
    for name in sp.FIELDS[1:]:
        fields[name] = bs.synth(boxk, name)[0]
    st = int(g["stride"])
    for name in sp.FIELDS:
        ref = g["box_" + name]
        got = fields[name].cpu().numpy().ravel()[::st]
        assert np.sqrt(((got - ref) ** 2).sum() / (ref ** 2).sum()) < 1e-5, name
    geom = sp.SkewerGeometry(NX, NY, NZ, dcell)
    q = np.concatenate(qso_files_from_golden(g))
    xyzr, nfor = sp.qso_lines_of_sight(geom, q["RA"], q["DEC"], q["Z_QSO_RSD"], float(g["ra0"]), float(g["dec0"]))
    eng = sp.SkewerEngine(geom, device=cuda)
    out = eng.read_spec(fields, xyzr, np.maximum(nfor, 0))
    _compare_with_pieces(g, geom, q, out)
    # merge stage: the oracle's bookkeeping (pinned against merge_spectra.py on CPU) replays the reference's
    # np.random stream per HDU and hands back the white noise each written forest consumed; delta_s and FGPA
    # then run on the GPU from the GPU skewers.
    ids = list(q["THING_ID"])
    pieces = []
    for key in [k[:-len("_THING_ID")] for k in g if k.startswith("spectra_") and k.endswith("_THING_ID")]:
        for r, ID in enumerate(g[key + "_THING_ID"]):
            m = g[key + "_LAMBDA"][r] > 0
            row = ids.index(ID)
            pieces.append(dict(id=int(ID), hdu=int(q["HDU"][row]), ra=q["RA"][row], dec=q["DEC"][row],
                               z=q["Z_QSO_RSD"][row], lam=g[key + "_LAMBDA"][r][m], delta_l=g[key + "_DELTA_L"][r][m],
                               eta_par=g[key + "_ETA_PAR"][r][m], velo_par=g[key + "_VELO_PAR"][r][m],
                               redshift=g[key + "_REDSHIFT"][r][m]))
    p1d = om.P1DMissing()
    for mode, zfix in (("merged_zfix", 2.4), ("merged_z", None)):
        fg = sp.FGPA(geom, zfix=zfix, device=cuda)
        merged = []
        for hdu in range(int(g["nslice"])):
            sel = [p for p in pieces if p["hdu"] == hdu]
            if sel:
                merged += om.merge_spectra_hdu(sel, hdu, int(g["seed"]), p1d, geom.npixeltot, zfix=zfix)
        by_id = {m["id"]: m for m in merged}
        ref_ids = list(g[mode + "_THING_ID"])
        rows = np.array([ids.index(i) for i in ref_ids])
        noise = np.array([by_id[i]["noise"] for i in ref_ids])
        nf_merge = fg.forest_count(q["Z_QSO_RSD"][rows])
        ds = fg.small_scales(nf_merge, noise=noise)
        F = fg.flux(out[0][rows].contiguous(), ds, out[1][rows].contiguous()).cpu().numpy()
        assert np.max(np.abs(ds.cpu().numpy() - g[mode + "_DELTA_S"])) < 1e-5
        assert np.max(np.abs(F - g[mode + "_FLUX"])) < 1e-5
    bs.close()


def test_smallscale_and_fgpa_vs_oracle(cuda):
    from oracle import merge as om
    from saclaymocks_b200 import spectra as sp
    geom = sp.SkewerGeometry(32, 32, 1536, 2.19)
    rng = np.random.default_rng(5)
    nq = 7
    zq = rng.uniform(2.0, 3.5, nq)
    for zfix in (2.4, None):
        fg = sp.FGPA(geom, zfix=zfix, device=cuda)
        nf = fg.forest_count(zq)
        noise = rng.normal(size=(nq, 8192))
        ds = fg.small_scales(nf, noise=noise).cpu().numpy()
        p1d = om.P1DMissing()
        z = fg.z
        for i in range(nq):
            zeff = z[:nf[i]].mean()
            ref = om.small_scale_field(noise[i], geom.npixeltot, zeff, z, p1d)
            assert np.max(np.abs(ds[i] - ref)) < 2e-5      # |delta_s| reaches ~10: 2e-6 relative
        dl = rng.normal(0, 1.2, (nq, geom.npixeltot)).astype(np.float32)
        dl[:, -100:] = -1e6
        eta = rng.normal(0, 0.5, (nq, geom.npixeltot)).astype(np.float32)
        F = fg.flux(torch.as_tensor(dl, device=cuda), torch.as_tensor(ds, device=cuda),
                    torch.as_tensor(eta, device=cuda)).cpu().numpy()
        ref = om.fgpa(dl.astype(np.float64) + ds, eta.astype(np.float64), fg.growthf, fg.a.cpu().numpy().astype(np.float64),
                      fg.b.cpu().numpy().astype(np.float64), fg.c.cpu().numpy().astype(np.float64))
        assert np.max(np.abs(F - ref)) < 1e-5
        assert np.all(F[:, -100:] == 1.0)


def test_register_blocked_kernel_equals_simple_kernel(cuda, golden_ref32):
    """The P-pixels-per-thread gather (union window) against the one-pixel-per-thread kernel on the same input."""
    from saclaymocks_b200 import spectra as sp
    g = golden_ref32
    rng = np.random.default_rng(3)
    boxes = {k: rng.standard_normal((32, 32, 1536), dtype=np.float32) for k in sp.FIELDS}
    from saclaymocks_b200 import _lib
    with _lib.option("skewers_kernel", 2):
        _, _, _, _, simple = _gpu_skewers(g, boxes, cuda, slabs=2)
    _, _, _, _, multi = _gpu_skewers(g, boxes, cuda, slabs=2)
    for a, b, tol in zip(simple, multi, (3e-6, 3e-6, 3e-6)):
        a, b = a.cpu().numpy(), b.cpu().numpy()
        assert np.array_equal(np.isnan(a), np.isnan(b))
        m = ~np.isnan(a)
        assert m.sum() > 10000
        assert np.max(np.abs(a[m] - b[m])) < tol


def test_staged_gather_equals_global_memory_gather(cuda):
    """The TMA-staged kernel (boxes of cells in shared memory) against the global-memory kernel on the same input:
    the same walk over the same values (agreement to the last bits: 2e-6; the two kernels instantiate the walk
    separately, so the compiler may contract a multiply-add differently); most segments must really be staged."""
    from saclaymocks_b200 import _lib
    from saclaymocks_b200 import spectra as sp
    rng = np.random.default_rng(3)
    NX, NY, NZ, dcell = 96, 64, 1536, 2.19
    geom = sp.SkewerGeometry(NX, NY, NZ, dcell)
    boxes = {k: torch.as_tensor(rng.standard_normal((NX, NY, NZ), dtype=np.float32), device=cuda) for k in sp.FIELDS}
    nq = 300
    hx = np.degrees(np.arctan(geom.LX / 2 / (geom.R0 + geom.LZ / 2))) * 1.05      # a few sightlines leave the box
    hy = np.degrees(np.arctan(geom.LY / 2 / (geom.R0 + geom.LZ / 2))) * 1.05
    ra = (190.0 + rng.uniform(-hx, hx, nq)).astype("f4")
    dec = rng.uniform(-hy, hy, nq).astype("f4")
    z = rng.uniform(1.9, 3.55, nq).astype("f4")
    xyzr, nfor = sp.qso_lines_of_sight(geom, ra, dec, z, 190.0, 0.0)
    eng = sp.SkewerEngine(geom, device=cuda)
    for rsd, dla in ((True, True), (True, False), (False, False)):
        staged = eng.read_spec(boxes, xyzr, nfor, rsd=rsd, dla=dla)
        seg, back, box = _lib.skewers_stats()
        assert seg > 0 and back < 0.25 * seg, (seg, back, box)                  # the staged kernel did the work
        with _lib.option("skewers_kernel", 1):
            plain = eng.read_spec(boxes, xyzr, nfor, rsd=rsd, dla=dla)
            assert _lib.skewers_stats()[0] == 0
        for a, b in zip(staged, plain):
            assert torch.equal(torch.isnan(a), torch.isnan(b))
            assert float((torch.nan_to_num(a, nan=-7.0) - torch.nan_to_num(b, nan=-7.0)).abs().max()) < 2e-6
        assert int((~torch.isnan(staged[0])).sum()) > 100000
    # two x-slabs with halo planes: the same rows again
    out = None
    for s_ in range(2):
        lo, hi = max(s_ * NX // 2 - 3, 0), min((s_ + 1) * NX // 2 + 3, NX)
        f = {k: boxes[k][lo:hi].contiguous() for k in sp.FIELDS}
        out = eng.read_spec(f, xyzr, nfor, ix0=lo, xmin=geom.LX * s_ / 2 - geom.LX / 2,
                            xmax=geom.LX * (s_ + 1) / 2 - geom.LX / 2, out=out)
    full = eng.read_spec(boxes, xyzr, nfor)
    for a, b in zip(out, full):
        assert torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))


def test_edge_cases_empty_and_out_of_range(cuda, golden_small):
    """Empty catalogue, quasars outside [zmin, zmax] (make_spectra.py:437-438), a sightline outside the slab."""
    from saclaymocks_b200 import spectra as sp
    g = golden_small
    geom = sp.SkewerGeometry(int(g["NX"]), int(g["NY"]), int(g["NZ"]), float(g["dcell"]))
    eng = sp.SkewerEngine(geom, device=cuda)
    f = {k: torch.as_tensor(g["box_" + k], device=cuda) for k in sp.FIELDS}
    out = eng.read_spec(f, np.zeros((0, 4)), np.zeros(0, dtype=np.int32))
    assert out[0].shape == (0, geom.npixeltot)
    xyzr, nfor = sp.qso_lines_of_sight(geom, np.float32([190.0, 190.1, 190.0]), np.float32([0.0, 0.1, 0.0]),
                                       np.float32([1.5, 2.5, 3.9]), 190.0, 0.0)
    assert list(nfor < 0) == [True, False, True]                      # z outside [1.8, 3.6] is dropped by the caller
    # a slab that the sightline never enters leaves every pixel untouched (NaN initialisation)
    d, e, v = eng.read_spec(f, xyzr[1:2], nfor[1:2], xmin=geom.LX / 2 - 1.0, xmax=geom.LX / 2)
    assert bool(torch.isnan(d).all())
    # forest length 0: owned pixels get the sentinels of make_spectra.py:99-101
    d, e, v = eng.read_spec(f, xyzr[1:2], np.int32([0]))
    assert bool((d == -1e6).all()) and bool((e == 0).all()) and bool((v == 0).all())
    fg = sp.FGPA(geom, zfix=2.4, device=cuda)
    F = fg.flux(d, None, e)
    assert bool((F == 1.0).all())                                     # exp(-a exp(b G (-1e6))) == 1


def test_null_box_is_reported(cuda):
    """make_boxes.py:100-105: an all-zero box raises."""
    from saclaymocks_b200.boxes import BoxSynth
    bs = BoxSynth(16, 16, 24, 8.0, device=cuda)
    boxk = bs.draw_grf_boxk(seed=3)
    zeros = torch.zeros((16, 16, 13), dtype=torch.float32, device=cuda)
    box, stats = bs.synth(boxk, "boxln_1", wtable=zeros)
    with pytest.raises(ValueError):
        bs.sigma(stats)
    W = {k: zeros.cpu().numpy() for k in ("Pln1", "Pln2", "Pln3", "P0")}
    from saclaymocks_b200 import _lib
    with pytest.raises(_lib.SmkError, match="null"):
        bs.make_boxes_host(W, seed=3)
    bs.close()
