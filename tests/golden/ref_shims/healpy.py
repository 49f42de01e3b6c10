"""Shim: healpy is absent; only ang2pix is reached on the hot path (util.radec2pix)."""
from saclaymocks_b200.healpix import ang2pix   # noqa


class pixelfunc(object):
    pass


def nside2npix(nside):
    """12 nside^2 (healpy.nside2npix), reached by bin/make_transmissions.py:35."""
    return 12 * nside * nside
