"""GRF box synthesis — restated from bin/make_boxes.py:40-125 (DrawGRF_boxk, FFTandStore),
:163-171 (seed), :239-431 (the 13 products).  pyfftw -> scipy.fft (float32, pocketfft)."""
import numpy as np
import scipy.fft as sfft

from . import cosmology as co

PRODUCTS = ("boxln_1", "boxln_2", "boxln_3", "box", "eta_xx", "eta_yy", "eta_zz", "eta_xy", "eta_xz", "eta_yz",
            "vx", "vy", "vz")


def draw_noise(NX, NY, NZ, seed):
    """make_boxes.py:46-48 + :163-171: legacy global MT19937 stream, one [NX,NY] plane per iz."""
    np.random.seed(seed)
    box = np.zeros((NX, NY, NZ), dtype=np.float32)
    for iz in range(NZ):
        box[:, :, iz] = np.float32(np.random.normal(size=[NX, NY]))
    return box


def draw_noise_planes(NX, NY, NZ, seed):
    """The same stream as draw_noise, returned plane-major [NZ][NX][NY] (contiguous writes: 14 s instead of a minute
    at 512 x 512 x 1536); draw_noise(...)[x, y, z] == draw_noise_planes(...)[z, x, y].  The GPU tests upload this
    and permute on the device."""
    np.random.seed(seed)
    box = np.empty((NZ, NX, NY), dtype=np.float32)
    for iz in range(NZ):
        box[iz] = np.random.normal(size=[NX, NY])
    return box


def forward(box, workers=1):
    """make_boxes.py:52-54: unnormalised r2c over all three axes, complex64."""
    boxk = sfft.rfftn(box, axes=(0, 1, 2), workers=workers)
    assert boxk.dtype == np.complex64
    return boxk


def backward(boxk, nz, workers=1):
    """make_boxes.py:86-92: unnormalised c2r, /= N, sigma = np.std(box)."""
    NX, NY = boxk.shape[:2]
    box = sfft.irfftn(boxk, s=(NX, NY, nz), axes=(0, 1, 2), workers=workers, norm="forward")
    box /= NX * NY * nz
    return box


def kgrid(NX, NY, NZ, dcell):
    """make_boxes.py:299-306."""
    k_ny = np.pi / dcell
    kx = np.fft.fftfreq(NX) * 2 * k_ny
    ky = np.fft.fftfreq(NY) * 2 * k_ny
    kz = np.fft.rfftfreq(NZ) * 2 * k_ny
    kz = np.float32(kz)
    ky = np.float32(ky.reshape(-1, 1))
    kx = np.float32(kx.reshape(-1, 1, 1))
    kk = kx * kx + ky * ky + kz * kz
    kk[0, 0, 0] = 1
    return kx, ky, kz, kk


def dgrowth0():
    """make_boxes.py:313."""
    return co.tables()["dgrowth_dDdz"][0]


def product_boxk(name, boxk_raw, boxk_p0, W, kg):
    """The k-space array make_boxes feeds to FFTandStore for product `name` (make_boxes.py:247-429)."""
    kx, ky, kz, kk = kg
    if name.startswith("boxln_"):
        return boxk_raw * W["Pln" + name[-1]]
    if name == "box":
        return boxk_p0.copy()
    if name.startswith("eta_"):
        ki = {"x": kx, "y": ky, "z": kz}
        b = boxk_p0.copy()
        b *= ki[name[4]] * ki[name[5]] / kk
        return b
    if name in ("vx", "vy", "vz"):
        ki = {"x": kx, "y": ky, "z": kz}[name[1]]
        b = boxk_p0.copy()
        b *= -1j * ki / kk * co.H0 * dgrowth0()
        return b
    raise KeyError(name)


def make_boxes(NX, NY, NZ, dcell, seed, W, workers=1, noise=None, products=PRODUCTS):
    """Returns (boxk_raw, boxk_p0, {name: float32 box}, {name: sigma})."""
    if noise is None:
        noise = draw_noise(NX, NY, NZ, seed)
    boxk_raw = forward(noise, workers)
    boxk_p0 = boxk_raw * W["P0"]                      # make_boxes.py:289-291
    kg = kgrid(NX, NY, NZ, dcell)
    boxes, sigmas = {}, {}
    for name in products:
        box = backward(product_boxk(name, boxk_raw, boxk_p0, W, kg), NZ, workers)
        boxes[name] = box
        sigmas[name] = np.std(box)
    return boxk_raw, boxk_p0, boxes, sigmas
