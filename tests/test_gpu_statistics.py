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
