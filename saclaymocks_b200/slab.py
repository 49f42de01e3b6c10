"""Host-side description of the slab decomposition: who owns what, and the buffer layouts on
both sides of the all-to-all.  chunk.py uses the ownership helpers; the layout helpers are the numpy statement of
the addressing the strided FFT kernels use (smk_capi.cu: PassAddr of smk_fft_r2c_local / smk_synth_c2r_finish) and
are exercised with a real collective (gloo) in tests/test_dist_cpu.py."""
import numpy as np


def x_planes(rank, nranks, NX):
    """Global x-planes [lo, hi) owned by `rank` (make_spectra.py:220-221 without the halo)."""
    return rank * NX // nranks, (rank + 1) * NX // nranks


def halo(rank, nranks, dmax):
    """Halo planes below / above the owned slab: global edges get none (make_spectra.py:220-221)."""
    return (dmax if rank > 0 else 0), (dmax if rank < nranks - 1 else 0)


def x_bounds(rank, nranks, LX):
    """Ownership interval xmin < X <= xmax of the slab (make_spectra.py:279-280, 443-448)."""
    return LX * rank / nranks - LX / 2, LX * (rank + 1) / nranks - LX / 2


def touching(xyzr, r_first, r_last, xmin, xmax):
    """Quasars whose sightline (X grows monotonically with R) can own a pixel of the slab."""
    xa = xyzr[:, 0] * r_first / xyzr[:, 3]
    xb = xyzr[:, 0] * r_last / xyzr[:, 3]
    return (np.minimum(xa, xb) <= xmax) & (np.maximum(xa, xb) > xmin)


def home_rank(xyzr, nranks, LX):
    """Rank whose slab contains the quasar itself (names the output file, like the reference's HDU index)."""
    return np.clip(np.floor((xyzr[:, 0] + LX / 2) / (LX / nranks)).astype(int), 0, nranks - 1)


# ---- all-to-all layouts (P = padded row length of boxk)
def forward_pack(local_xyk, nranks):
    """[nxl][NY][P] after the local z and y passes -> send buffer [dest][nxl][nyl][P] (dest = y // nyl)."""
    nxl, NY, P = local_xyk.shape
    nyl = NY // nranks
    return np.ascontiguousarray(local_xyk.reshape(nxl, nranks, nyl, P).transpose(1, 0, 2, 3))


def forward_unpack(recv, nranks):
    """Received [src][nxl][nyl][P] IS boxk's y-slab layout [NX][nyl][P] (x = src*nxl + xl): a reshape."""
    src, nxl, nyl, P = recv.shape
    return recv.reshape(src * nxl, nyl, P)


def inverse_pack(boxk_yslab, nranks):
    """[NX][nyl][P] after the x pass: destination chunks (x ranges) are already contiguous."""
    NX, nyl, P = boxk_yslab.shape
    return boxk_yslab.reshape(nranks, NX // nranks, nyl, P)


def inverse_unpack(recv):
    """Received [src][nxl][nyl][P] -> the x-slab [nxl][NY][P] the y pass reads (y = src*nyl + yl), done in the
    kernel by two-level addressing."""
    src, nxl, nyl, P = recv.shape
    return np.ascontiguousarray(recv.transpose(1, 0, 2, 3)).reshape(nxl, src * nyl, P)
