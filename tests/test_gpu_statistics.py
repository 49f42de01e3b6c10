"""Statistical acceptance with GPU Philox noise (BASELINE.json north_star): the 3-D power spectrum of the delta box
and the 1-D power spectrum of the small-scale field agree with the input spectra within the mode-count error.
torch.fft is used here as an independent cross-check only (never in the product path)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def test_recovered_pk_matches_input(cuda):
    """P(k) of the delta box made from Philox noise vs the input P0 (powerspectrum.P_0): <|delta_k|^2> V / N^2."""
    from saclaymocks_b200 import pk
    from saclaymocks_b200.boxes import BoxSynth
    NX, NY, NZ, dcell = 128, 128, 256, 2.19
    bs = BoxSynth(NX, NY, NZ, dcell, device=cuda)
    W = pk.weight_tables(NX, NY, NZ, dcell)
    boxk = bs.draw_grf_boxk(seed=2024)
    box, stats = bs.synth(boxk, "box", wtable=bs.upload_weights(W["P0"]))
    # independent estimator
    dk = torch.fft.rfftn(box.double())
    N = NX * NY * NZ
    vol = N * dcell ** 3
    p3d = (dk.abs() ** 2 * vol / N ** 2).cpu().numpy()
    k = pk.k_norm(NX, NY, NZ, dcell).astype(np.float64)
    k_ny = np.pi / dcell
    edges = np.linspace(0.05, 0.9 * k_ny, 31)
    # Hermitian multiplicity: planes kz=0 and kz=Nyquist count once, the others twice
    mult = np.full(NZ // 2 + 1, 2.0)
    mult[0] = mult[-1] = 1.0
    mult = np.broadcast_to(mult, k.shape)
    idx = np.digitize(k.ravel(), edges)
    chi2 = []
    for b in range(1, len(edges)):
        m = idx == b
        nm = mult.ravel()[m].sum() / 2.0                   # independent complex modes
        est = (p3d.ravel()[m] * mult.ravel()[m]).sum() / mult.ravel()[m].sum()
        ref = (np.maximum(pk.spline("P0")(k.ravel()[m]), 0) * mult.ravel()[m]).sum() / mult.ravel()[m].sum()
        chi2.append(((est / ref - 1) / np.sqrt(1.0 / nm)) ** 2)
    chi2 = np.array(chi2)
    assert chi2.max() < 25.0, chi2                          # no bin beyond 5 sigma of the mode-count error
    assert chi2.mean() < 2.0, chi2.mean()                   # and the ensemble is consistent with unit chi^2
    # sigma of the box against the prediction from the weights (SURVEY Appendix B anchor for C1: 2.48)
    w2 = (W["P0"].astype(np.float64) ** 2 * mult).sum() / N
    assert abs(bs.sigma(stats) / np.sqrt(w2) - 1) < 5e-3
    bs.close()


def test_small_scale_p1d_matches_input(cuda):
    """P1D of delta_s drawn with Philox vs P1D_missing(z_eff, k) (merge_spectra.py:308-324; estimator of
    powerspectrum.P1D_1spectrum: |rfft|^2 * pix / n)."""
    from saclaymocks_b200 import spectra as sp
    geom = sp.SkewerGeometry(32, 32, 1536, 2.19)
    fg = sp.FGPA(geom, zfix=2.4, device=cuda)
    nq = 4000
    zq = np.full(nq, 3.5)
    nf = fg.forest_count(zq)
    ds = fg.small_scales(nf, seed=99, qso_ids=np.arange(nq))
    npix, pix = geom.npixeltot, 0.2
    spec = torch.fft.rfft(ds.double(), dim=1)
    p1d = (spec.abs() ** 2 * pix / npix).mean(dim=0).cpu().numpy()
    k = np.fft.rfftfreq(npix) * 2 * np.pi / pix
    iz = fg.p1d.iz(fg.zeff(nf)[0])
    # expectation of that estimator for the reference's procedure: a periodic field of nfft samples with spectrum
    # P_miss/pix, of which the first npix samples are kept: E|X_j|^2 = sum_tau (npix-|tau|) C[tau] exp(-2 pi i j tau/npix)
    nfft = fg.nfft_for(npix)
    kf = np.fft.rfftfreq(nfft) * 2 * np.pi / pix
    corr = np.fft.irfft(np.maximum(np.interp(kf, fg.p1d.k, fg.p1d.pk[iz]), 0) / pix, n=nfft)
    g = (npix - np.arange(npix)) * corr[:npix]
    ref = (2 * np.fft.fft(g).real - g[0])[:npix // 2 + 1] * pix / npix
    edges = np.linspace(0.2, 12.0, 25)
    for a, b in zip(edges[:-1], edges[1:]):
        m = (k >= a) & (k < b)
        nm = m.sum() * nq
        assert abs(p1d[m].mean() / ref[m].mean() - 1) < 5 / np.sqrt(nm), (a, b)
    # and the windowed expectation itself stays within a few per cent of the input P1D_missing
    raw = np.interp(k, fg.p1d.k, fg.p1d.pk[iz])
    sel = (k > 0.2) & (k < 12) & (raw > 0.02)        # P1D_missing crosses zero at high k (clipped at 0 by the reference)
    assert np.max(np.abs(ref[sel] / raw[sel] - 1)) < 0.05
    # different quasar ids give independent draws, same id reproduces
    ds2 = fg.small_scales(nf[:4], seed=99, qso_ids=np.arange(4))
    assert torch.equal(ds2, ds[:4])
    assert float((ds[0] * ds[1]).mean().abs()) < 0.2 * float((ds[0] ** 2).mean())


def test_gpu_pk_estimator_matches_independent_estimator(cuda):
    """smk_pk_estimate (own r2c + binning kernel) against the torch.fft / numpy estimator above, bin by bin, and
    against the input P0 within the mode-count error."""
    from saclaymocks_b200 import pk
    from saclaymocks_b200.boxes import BoxSynth
    NX, NY, NZ, dcell = 128, 64, 256, 2.19
    bs = BoxSynth(NX, NY, NZ, dcell, device=cuda)
    W = pk.weight_tables(NX, NY, NZ, dcell)
    boxk = bs.draw_grf_boxk(seed=7)
    box, _ = bs.synth(boxk, "box", wtable=bs.upload_weights(W["P0"]))
    nb, kmin, kmax = 24, 0.05, 0.9 * np.pi / dcell
    kmean, P, nmodes = bs.power_spectrum(box, nbins=nb, kmin=kmin, kmax=kmax)
    dk = torch.fft.rfftn(box.double())
    N = NX * NY * NZ
    p3d = (dk.abs() ** 2 * (N * dcell ** 3) / N ** 2).cpu().numpy().ravel()
    k = pk.k_norm(NX, NY, NZ, dcell).astype(np.float32).ravel()
    mult = np.full(NZ // 2 + 1, 2.0)
    mult[0] = mult[-1] = 1.0
    mult = np.broadcast_to(mult, (NX, NY, NZ // 2 + 1)).ravel()
    b = np.floor((k - np.float32(kmin)) * np.float32(nb / (kmax - kmin))).astype(int)
    for i in range(nb):
        m = b == i
        assert abs(nmodes[i] - mult[m].sum() / 2) <= 2              # float32 bin edges: at most a couple of modes move
        ref = (p3d[m] * mult[m]).sum() / mult[m].sum()
        assert abs(P[i] / ref - 1) < 2e-3, (i, P[i], ref)
        assert abs(kmean[i] - (k[m] * mult[m]).sum() / mult[m].sum()) < 1e-4
        truth = (np.maximum(pk.spline("P0")(k[m].astype(np.float64)), 0) * mult[m]).sum() / mult[m].sum()
        assert abs(P[i] / truth - 1) < 5.5 / np.sqrt(nmodes[i]), i
    bs.close()


def test_gpu_p1d_estimator_matches_numpy(cuda):
    """smk_p1d (GPU P1D_1spectrum averaged like ComputeP1D, powerspectrum.py:204-238) against numpy's rfft on the same
    windows: plain rows, windows with an offset / too-short rows skipped, and the contrast F / <F> - 1."""
    from saclaymocks_b200 import spectra as sp
    rng = np.random.default_rng(11)
    nq, npix, pix = 64, 6524, 0.2
    rows = rng.standard_normal((nq, npix)).astype(np.float32)
    rows_d = torch.as_tensor(rows, device=cuda)
    for nfft in (256, 2048, 4096):
        first = rng.integers(0, 200, nq).astype(np.int32)
        nvalid = rng.integers(nfft - 300, npix - 200, nq).astype(np.int32)
        mean = (1.0 + 0.2 * rng.random(nq)).astype(np.float32)
        est = sp.P1DEstimator(nfft, pix, device=cuda)
        est.add(rows_d, first=first, nvalid=nvalid, mean=mean)
        k, p1d, err, n = est.result()
        use = [q for q in range(nq) if nvalid[q] >= nfft]
        assert n == len(use) and 0 < n < nq
        ref = np.array([np.abs(np.fft.rfft(rows[q, first[q]:first[q] + nfft].astype(np.float64) / mean[q] - 1)) ** 2
                        * pix / nfft for q in use])
        assert np.allclose(k, np.fft.rfftfreq(nfft) * 2 * np.pi / pix)
        assert np.max(np.abs(p1d / ref.mean(axis=0) - 1)) < 2e-5
        assert np.max(np.abs(err / (ref.std(axis=0) / np.sqrt(n)) - 1)) < 1e-3
    est = sp.P1DEstimator(4096, pix, device=cuda)          # two calls accumulate
    est.add(rows_d[:30])
    est.add(rows_d[30:])
    assert est.result()[3] == nq
    ref = (np.abs(np.fft.rfft(rows[:, :4096].astype(np.float64), axis=1)) ** 2 * pix / 4096).mean(axis=0)
    assert np.max(np.abs(est.result()[1] / ref - 1)) < 2e-5


def test_p1d_of_flux_against_cpu_oracle_on_the_same_boxes(cuda):
    """Acceptance of the end product: P1D of delta_F = F / <F> - 1 from the GPU chain (Philox boxes -> skewers ->
    delta_s -> FGPA -> smk_p1d) against the CPU oracle's chain run on the SAME boxes and the same small-scale noise
    (oracle ReadSpec + small-scale field + fgpa, numpy P1D_1spectrum): the two P1D agree to 1e-4 per wavenumber,
    far inside the mode-count error of either, and the mean transmissions to 1e-6."""
    from oracle import merge as om
    from oracle import spectra as osp
    from saclaymocks_b200 import pk
    from saclaymocks_b200 import spectra as sp
    from saclaymocks_b200.boxes import BoxSynth, WEIGHT_OF
    NX, NY, NZ, dcell = 32, 32, 1536, 2.19
    bs = BoxSynth(NX, NY, NZ, dcell, device=cuda)
    W = pk.weight_tables(NX, NY, NZ, dcell)
    boxk = bs.draw_grf_boxk(seed=17)
    fields = {"box": bs.synth(boxk, "box", wtable=bs.upload_weights(W["P0"]))[0]}
    for name in sp.FIELDS[1:]:
        fields[name] = bs.synth(boxk, name)[0]
    geom = sp.SkewerGeometry(NX, NY, NZ, dcell)
    rng = np.random.default_rng(3)
    nq = 48
    half = np.degrees(np.arctan((geom.LX / 2 - 4 * dcell) / (geom.R0 + geom.LZ / 2))) * 0.9
    q = np.zeros(nq, dtype=[("RA", "f4"), ("DEC", "f4"), ("Z_QSO_NO_RSD", "f4"), ("Z_QSO_RSD", "f4"),
                            ("THING_ID", "i8"), ("HDU", "i4")])
    q["RA"], q["DEC"] = 190.0 + rng.uniform(-half, half, nq), rng.uniform(-half, half, nq)
    q["Z_QSO_RSD"] = q["Z_QSO_NO_RSD"] = rng.uniform(3.2, 3.55, nq)
    q["THING_ID"] = np.arange(nq)
    xyzr, nfor = sp.qso_lines_of_sight(geom, q["RA"], q["DEC"], q["Z_QSO_RSD"], 190.0, 0.0)
    eng = sp.SkewerEngine(geom, device=cuda)
    dl, ep, vp = eng.read_spec(fields, xyzr, nfor)
    fg = sp.FGPA(geom, zfix=2.4, device=cuda)
    nfm = fg.forest_count(q["Z_QSO_RSD"])
    noise = rng.normal(size=(nq, fg.nfft_for(geom.npixeltot)))
    ds = fg.small_scales(nfm, noise=noise)
    F = fg.flux(dl, ds, ep)
    nfft = 4096
    est = sp.P1DEstimator(nfft, 0.2, device=cuda)
    meanF = torch.stack([F[i, :nfft].mean() for i in range(nq)])
    est.add(F, nvalid=nfm.astype(np.int32), mean=meanF)
    k, p_gpu, err, n = est.result()
    assert n == nq                                                   # every forest is longer than the window
    # ---- the CPU oracle on the same boxes
    boxes = {kf: v.cpu().numpy() for kf, v in fields.items()}
    og = osp.Geometry(NX, NY, NZ, dcell)
    lam32 = np.float32(geom.lambda_vec)
    p1d = om.P1DMissing()
    z32 = fg.z
    a, b, c = (getattr(fg, x).cpu().numpy().astype(np.float64) for x in "abc")
    p_cpu, mean_diff = [], 0.0
    pieces = {pc["id"]: pc for pc in osp.make_spectra_slice(og, boxes, [q], 0, 1, 190.0, 0.0)}
    assert len(pieces) == nq
    for i in range(nq):
        pc = pieces[i]
        idx = np.searchsorted(lam32, pc["lam"])
        assert idx[0] == 0 and len(idx) >= nfft
        zeff = z32[:nfm[i]].mean()
        dsc = om.small_scale_field(noise[i], geom.npixeltot, zeff, z32, p1d)
        Fc = om.fgpa(np.float64(pc["delta_l"][:nfft]) + dsc[:nfft], np.float64(pc["eta_par"][:nfft]), fg.growthf[:nfft],
                     a[:nfft], b[:nfft], c[:nfft])
        mean_diff = max(mean_diff, abs(Fc.mean() - float(meanF[i])))
        p_cpu.append(np.abs(np.fft.rfft(Fc / Fc.mean() - 1)) ** 2 * 0.2 / nfft)
    p_cpu = np.mean(p_cpu, axis=0)
    assert mean_diff < 1e-6
    sel = (k > 0.05) & (k < 10.0)
    assert np.max(np.abs(p_gpu[sel] / p_cpu[sel] - 1)) < 1e-4
    assert np.all(err[sel] / p_gpu[sel] > 5e-2)                      # the statistical error is orders of magnitude larger
    bs.close()


def test_gpu_lognormal_pk_matches_host(cuda):
    """GPU LogNormalP (powerspectrum.py:194-200 with the two 1-D float64 transforms on the GPU, smk_fft1d_f64) against
    the host version (numpy / pocketfft): the transform itself to 1e-12 of the largest mode, P_ln(k) to 1e-9 relative
    wherever it matters for the weight tables."""
    from saclaymocks_b200 import pk
    fft = pk.gpu_fft(cuda)
    rng = np.random.default_rng(2)
    for n in (2, 8, 1 << 10, 1 << 19, 1 << 20):
        a = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        ref = np.fft.fft(a)
        assert np.max(np.abs(fft(a) - ref)) < 1e-12 * max(1.0, np.abs(ref).max()), n
    k, P = pk._input_pk(pk.fgrowth(2.75, pk.constant.omega_M_0) * pk.bias_qso(2.75))
    kh, Ph = pk.lognormal_pk(k, P)
    kg, Pg = pk.lognormal_pk(k, P, fft=fft)
    assert np.array_equal(kh, kg)
    sel = Ph > 1e-6 * Ph.max()
    assert np.max(np.abs(Pg[sel] / Ph[sel] - 1)) < 1e-9
