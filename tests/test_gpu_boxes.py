"""GPU parity of the box synthesis (libsmk.so through the C ABI) against the CPU oracle and the reference goldens.
Tolerance: 1e-5 relative L2 on every box (BASELINE.json north_star)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import rel_l2  # noqa: E402

TOL = 1e-5


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _oracle_run(NX, NY, NZ, dcell, seed, W=None):
    from oracle import boxes as ob
    from oracle import pk_weights
    if W is None:
        W = pk_weights.weights(NX, NY, NZ, dcell)
    noise = ob.draw_noise(NX, NY, NZ, seed)
    raw, p0, boxes, sig = ob.make_boxes(NX, NY, NZ, dcell, seed, W, workers=8, noise=noise)
    return W, noise, raw, p0, boxes, sig


@pytest.mark.parametrize("shape,dcell", [((16, 16, 24), 8.0), ((16, 16, 96), 35.04), ((32, 64, 96), 4.0),
                                         ((64, 32, 64), 2.19), ((128, 128, 256), 2.19), ((32, 32, 1536), 2.19)])
def test_boxes_match_oracle(cuda, shape, dcell):
    from saclaymocks_b200.boxes import BoxSynth, PRODUCTS, WEIGHT_OF
    NX, NY, NZ = shape
    W, noise, raw, p0, boxes, sig = _oracle_run(NX, NY, NZ, dcell, 42)
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


@pytest.mark.parametrize("group,streams", [(1, 1), (3, 2), (8, 3), (64, 2)])
def test_boxes_chained_yz_match_oracle(cuda, group, streams):
    """y<->z passes chained through an L2-resident scratch slot (plane groups on internal streams, ragged last
    group, L2 discard of the dead slot lines): same results as the oracle, forward and inverse."""
    from saclaymocks_b200 import _lib
    from saclaymocks_b200.boxes import BoxSynth, PRODUCTS, WEIGHT_OF
    NX, NY, NZ, dcell = 32, 64, 96, 4.0
    W, noise, raw, p0, boxes, sig = _oracle_run(NX, NY, NZ, dcell, 42)
    with _lib.option("yz_group", group), _lib.option("yz_streams", streams):      # read when the context is created
        bs = BoxSynth(NX, NY, NZ, dcell, device=cuda)
    boxk = bs.draw_grf_boxk(noise=torch.as_tensor(noise, device=cuda))
    assert rel_l2(bs.boxk_to_numpy(boxk), raw) < TOL
    Wd = {k: bs.upload_weights(v) for k, v in W.items()}
    for name in PRODUCTS:
        box, stats = bs.synth(boxk, name, wtable=Wd.get(WEIGHT_OF.get(name)))
        assert rel_l2(box.cpu().numpy(), boxes[name]) < TOL, name
        assert abs(bs.sigma(stats) / sig[name] - 1) < 1e-4, name
    bs.close()


def test_boxes_match_reference_golden(cuda, golden_small):
    """16 x 16 x 96 run of the unmodified reference make_boxes.py (tests/golden/ref_small.npz)."""
    from oracle import boxes as ob
    from saclaymocks_b200.boxes import BoxSynth, PRODUCTS, WEIGHT_OF
    g = golden_small
    NX, NY, NZ, dcell = int(g["NX"]), int(g["NY"]), int(g["NZ"]), float(g["dcell"])
    noise = ob.draw_noise(NX, NY, NZ, int(g["seed"]))
    bs = BoxSynth(NX, NY, NZ, dcell, device=cuda)
    boxk = bs.draw_grf_boxk(noise=torch.as_tensor(noise, device=cuda))
    Wd = {k: bs.upload_weights(g["W_" + k]) for k in ("Pln1", "Pln2", "Pln3", "P0")}
    for name in PRODUCTS:
        box, stats = bs.synth(boxk, name, wtable=Wd.get(WEIGHT_OF.get(name)))
        assert rel_l2(box.cpu().numpy(), g["box_" + name]) < TOL, name
        assert abs(bs.sigma(stats) / float(g["sigma_" + name]) - 1) < 1e-4, name
    assert rel_l2(bs.boxk_to_numpy(boxk), g["boxkP0"]) < TOL
    bs.close()


def test_host_buffer_chain(cuda, golden_small):
    """smk_make_boxes_host: the single C-ABI call a make_boxes.py replacement makes."""
    from oracle import boxes as ob
    from saclaymocks_b200.boxes import BoxSynth, PRODUCTS
    g = golden_small
    NX, NY, NZ, dcell = int(g["NX"]), int(g["NY"]), int(g["NZ"]), float(g["dcell"])
    noise = ob.draw_noise(NX, NY, NZ, int(g["seed"]))
    bs = BoxSynth(NX, NY, NZ, dcell, device=cuda)
    W = {k: g["W_" + k] for k in ("Pln1", "Pln2", "Pln3", "P0")}
    boxes, sig = bs.make_boxes_host(W, noise_host=noise)
    for name in PRODUCTS:
        assert rel_l2(boxes[name], g["box_" + name]) < TOL, name
        assert abs(sig[name] / float(g["sigma_" + name]) - 1) < 1e-4, name
    bs.close()


def test_roundtrip_large(cuda):
    """Size-independent property at a BASELINE size: c2r(r2c(x)) / N == x with unit weights (256 x 256 x 1536)."""
    from saclaymocks_b200.boxes import BoxSynth
    NX, NY, NZ = 256, 256, 1536
    bs = BoxSynth(NX, NY, NZ, 2.19, device=cuda)
    x = bs.noise_philox(7)
    boxk = bs.draw_grf_boxk(noise=x)
    ones = torch.ones((NX, NY, NZ // 2 + 1), dtype=torch.float32, device=cuda)
    y, stats = bs.synth(boxk, "boxln_1", wtable=ones)
    err = float(torch.linalg.vector_norm((y - x).double()) / torch.linalg.vector_norm(x.double()))
    assert err < 2e-6, err
    # Parseval: sum |boxk|^2 with Hermitian multiplicity == N * sum x^2
    k2 = (boxk[:, :, :NZ // 2 + 1].abs().double() ** 2)
    tot = 2 * k2.sum() - k2[:, :, 0].sum() - k2[:, :, NZ // 2].sum()
    ref = (x.double() ** 2).sum() * NX * NY * NZ
    assert abs(float(tot / ref) - 1) < 1e-5
    bs.close()


def test_fused_philox_equals_unfused(cuda):
    from saclaymocks_b200.boxes import BoxSynth
    bs = BoxSynth(64, 64, 96, 2.19, device=cuda)
    x = bs.noise_philox(123)
    a = bs.draw_grf_boxk(noise=x)
    b = bs.draw_grf_boxk(seed=123)
    assert torch.equal(a, b)
    v = x.double()
    assert abs(float(v.mean())) < 5 / np.sqrt(v.numel())
    assert abs(float(v.var()) - 1) < 5 * np.sqrt(2 / v.numel())
    assert abs(float((v ** 4).mean()) - 3) < 0.1
    bs.close()


def test_philox_noise_matches_host_restatement(cuda):
    """smk_noise_philox against a numpy restatement of Philox4x32-10 (itself pinned to the Random123 known-answer
    vectors, tests/test_philox_cpu.py) + Box-Muller.  The device uses __logf / __sincosf: 2e-5 absolute."""
    from helpers import philox_normals_np
    from saclaymocks_b200.boxes import BoxSynth
    bs = BoxSynth(16, 16, 24, 8.0, device=cuda)
    for seed in (0, 123, (7 << 32) + 5):
        got = bs.noise_philox(seed).cpu().numpy().ravel()
        ref = philox_normals_np(seed, got.size)
        assert np.max(np.abs(got - ref)) < 2e-5, seed
    bs.close()


def test_unsupported_shape_fails_loudly(cuda):
    from saclaymocks_b200 import _lib
    from saclaymocks_b200.boxes import BoxSynth
    with pytest.raises(_lib.SmkError):
        BoxSynth(24, 24, 24, 2.19, device=cuda)


@pytest.mark.parametrize("shape,dcell", [((16, 16, 96), 35.04), ((64, 32, 64), 2.19)])
def test_gpu_weight_tables_match_interpolate_pk(cuda, shape, dcell, golden_small):
    """smk_pk_weights (GPU interpolate_pk) against the host spline tables / the reference's P-file."""
    from saclaymocks_b200 import pk
    from saclaymocks_b200.boxes import BoxSynth
    NX, NY, NZ = shape
    bs = BoxSynth(NX, NY, NZ, dcell, device=cuda)
    ref = pk.weight_tables(NX, NY, NZ, dcell)
    for name in ("Pln1", "Pln2", "Pln3", "P0"):
        got = bs.weight_table(name).cpu().numpy()
        r = ref[name]
        if shape == (16, 16, 96):
            r = golden_small["W_" + name]                      # written by the unmodified interpolate_pk.py
        same = (got == r).mean()
        assert same > 0.999, (name, same)                      # float64 power basis vs FITPACK B-splines: rare 1-ulp ties
        assert np.max(np.abs(got - r) / np.maximum(np.abs(r), 1e-30)) < 2.5e-7, name
    bs.close()
